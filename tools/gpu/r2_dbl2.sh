# partitioned rank writes in the doubling rounds: tests (small forced, cfg4 at a tenth), then the cfg4 line and its trace
TAG=${1:-r2h}
set -x
timeout 1200 python -m pytest tests/test_gpu_msd.py tests/test_gpu_single_block.py tests/test_gpu_blocks.py tests/test_gpu_shard.py -m gpu -q -x --tb=short 2>&1 | tail -6 | cut -c1-800
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --tb=short -k "tenth" 2>&1 | tail -4 | cut -c1-800
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-file-level --e2e-steps 1 2> gpurun_out/${TAG}_bench_cfg4_n1.err | grep "^{" > gpurun_out/${TAG}_bench_cfg4_n1.json
tail -c 600 gpurun_out/${TAG}_bench_cfg4_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg4_n1.json"))
print("cfg4 ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
print(d["kernels_ms_per_step"]); print(d["counters"])
PY
B3M_TRACE=1 timeout 600 python tools/profile_step.py --workload cfg4 --scale 1.0 --steps 1 2>&1 | grep "^\[T\]" > gpurun_out/${TAG}_cfg4_trace.txt
cat gpurun_out/${TAG}_cfg4_trace.txt | head -60
