set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu_validate.log 2>&1
tail -8 gpurun_out/pytest_gpu_validate.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_cfg3_validate.err | grep "^{" > gpurun_out/bench_cfg3_validate.json
tail -c 300 gpurun_out/bench_cfg3_validate.err
