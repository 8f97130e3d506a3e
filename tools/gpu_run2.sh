set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3_r01c.json 2> gpurun_out/bench_cfg3_r01c.err
tail -c 1500 gpurun_out/bench_cfg3_r01c.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01c.json 2>&1
cat gpurun_out/bench_ref_r01c.json | head -c 1200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py --workload cfg3 --scale 0.25 2>&1 | grep -v Warning | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_cfg3_n2_r01c.json 2> gpurun_out/bench_cfg3_n2_r01c.err
tail -c 1500 gpurun_out/bench_cfg3_n2_r01c.err
