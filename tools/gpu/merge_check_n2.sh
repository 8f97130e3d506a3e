set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py --nsym 400003 --itype pacterm --strategy merge --local-blocks 2 2>&1 | grep -v Warning | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dist_check.py --workload cfg4 --scale 0.02 2>&1 | grep -v Warning | tail -8
