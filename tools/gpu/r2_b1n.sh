# temporary: level-1 bits of a position-sharded build on N GPUs
N=${1:-4}; TAG=${2:-r2j}
set -x
for B1 in 9 8; do
B3M_TUNE_B1=$B1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 2 2> gpurun_out/${TAG}_b$B1.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n${N}_b$B1.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n${N}_b$B1.json"))
print("b1=$B1 N=$N ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "parity", d.get("parity_check",{}).get("ok"))
print(d["kernels_ms_per_step_max_over_ranks"]); print(d["build_timeline_ms"])
PY
done
