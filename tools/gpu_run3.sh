set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/bench_cfg3_r01r.err | grep "^{" > gpurun_out/bench_cfg3_r01r.json
tail -c 300 gpurun_out/bench_cfg3_r01r.err
timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu 2> gpurun_out/bench_cfg5_r01r.err | grep "^{" > gpurun_out/bench_cfg5_r01r.json
