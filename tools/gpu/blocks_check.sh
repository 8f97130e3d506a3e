set -x
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_golden.py tests/test_gpu_files.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload cfg3 --scale 0.25 --numblocks 4 --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/bench_cfg3q_nb4_r02i.json
