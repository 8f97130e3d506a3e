set -x
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_dist.py -x -q 2>&1 | tail -40 | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_cfg3_n2_r02g.err | grep "^{" > gpurun_out/bench_cfg3_n2_r02g.json
tail -c 800 gpurun_out/bench_cfg3_n2_r02g.err
