set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/bench_cfg3_r01y.err | grep "^{" > gpurun_out/bench_cfg3_r01y.json
tail -c 300 gpurun_out/bench_cfg3_r01y.err
