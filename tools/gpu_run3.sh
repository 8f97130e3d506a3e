set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg3_r01d.json 2> gpurun_out/bench_cfg3_r01d.err
tail -c 600 gpurun_out/bench_cfg3_r01d.err
cat gpurun_out/bench_cfg3_r01d.json
