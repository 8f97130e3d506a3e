set -x
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -25 | cut -c1-400
for w in cfg2 cfg5 cfg4; do
timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu 2> gpurun_out/bench_${w}_r01p.err | grep "^{" > gpurun_out/bench_${w}_r01p.json
tail -c 300 gpurun_out/bench_${w}_r01p.err
done
