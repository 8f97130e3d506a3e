set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/bench_cfg3_n8_r02g.err | grep "^{" > gpurun_out/bench_cfg3_n8_r02g.json
tail -c 1500 gpurun_out/bench_cfg3_n8_r02g.err
