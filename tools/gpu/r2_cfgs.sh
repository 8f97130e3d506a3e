# the other configs' lines on the final tree (short: no CPU leg, no file-level leg)
TAG=${1:-r2z}
for W in cfg1 cfg2 cfg5; do
  timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --no-file-level --e2e-steps 3 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_${W}_n1.json
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${W}_n1.json"))
print("$W ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],3), d["kernels_ms_per_step"])
PY
done
