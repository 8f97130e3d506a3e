set -x
timeout 600 python -m pytest tests/test_gpu_single_block.py tests/test_gpu_blocks.py -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/bench_cfg3_r02f.err | grep "^{" > gpurun_out/bench_cfg3_r02f.json
tail -c 300 gpurun_out/bench_cfg3_r02f.err
