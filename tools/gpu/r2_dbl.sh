# round 2: the prefix-doubling rounds with the CTA-local split (k_dbl_tile): tests, then cfg4 and cfg3 lines
TAG=${1:-r2d}
set -x
timeout 1200 python -m pytest tests/test_gpu_msd.py tests/test_gpu_single_block.py tests/test_gpu_blocks.py tests/test_gpu_multi_abi.py tests/test_gpu_memory.py -m gpu -q -x --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log | cut -c1-800
for WL in cfg4 cfg3; do
timeout 900 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --e2e-steps 1 2> gpurun_out/${TAG}_bench_${WL}_n1.err | grep "^{" > gpurun_out/${TAG}_bench_${WL}_n1.json
tail -c 600 gpurun_out/${TAG}_bench_${WL}_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${WL}_n1.json"))
print("$WL ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
print(d["phases_ms"]); print(d["kernels_ms_per_step"]); print(d["counters"])
PY
done
B3M_TRACE=1 timeout 600 python tools/profile_step.py --workload cfg4 --scale 1.0 --steps 1 2>&1 | grep "^\[T\]" > gpurun_out/${TAG}_cfg4_trace.txt
cat gpurun_out/${TAG}_cfg4_trace.txt | head -80
