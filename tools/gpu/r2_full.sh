# the whole GPU suite, smoke, and the default bench line (no profiling)
TAG=${1:-r2p}
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x --durations=6 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -25 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-600
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/${TAG}_bench_cfg3_n1.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n1.json
tail -c 400 gpurun_out/${TAG}_bench_cfg3_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n1.json"))
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roof", d["roofline"])
print(d["phases_ms"]); print(d["kernels_ms_per_step"]); print(d.get("file_level")); print(d.get("k8_rl_encode"))
PY
