# base / v1 / v2 libraries on cfg4 at a tenth (doubling kernels), v2 also on cfg3
TAG=${1:-r2w}
cp bwtb3m_b200/libb3m.so /tmp/base.so
for V in base v1 v2; do
if [ $V = base ]; then cp /tmp/base.so bwtb3m_b200/libb3m.so; else cp bwtb3m_b200/libb3m_$V.so.bin bwtb3m_b200/libb3m.so; fi
timeout 300 python bench.py --workload cfg4 --scale 0.1 --steps 3 --warmup 3 --no-cpu --no-file-level --e2e-steps 1 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_cfg4t_$V.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg4t_$V.json"))
k=d["kernels_ms_per_step"]
print("$V cfg4/10 ms/step", round(d["ms_per_step"],2), {x: k[x] for x in k if x.startswith("dbl") or x.startswith("heads") or x.startswith("radix") or x.startswith("extract")})
PY
done
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-file-level --e2e-steps 2 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_cfg3_v2.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_v2.json"))
print("v2 cfg3 ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["kernels_ms_per_step"])
PY
cp /tmp/base.so bwtb3m_b200/libb3m.so
