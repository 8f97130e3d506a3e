set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg3_r01o.json 2> gpurun_out/bench_cfg3_r01o.err
tail -c 600 gpurun_out/bench_cfg3_r01o.err
cat gpurun_out/bench_cfg3_r01o.json
