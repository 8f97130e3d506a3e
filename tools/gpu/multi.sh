# usage: gpurun --gpus N -- 'bash tools/gpu/multi.sh N TAG'   (N = 2, 4, 8)
# NCCL parity tests on all N GPUs (multi-GPU result == single-GPU result, incl. cfg3 at full size), then the N-GPU bench line
N=${1:-2}; TAG=${2:-r2}
set -x
timeout 1500 python -m pytest tests/test_gpu_dist.py -x -q --timeout 900 2>&1 | tail -15 | cut -c1-800
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_cfg3_n$N.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n$N.json
tail -c 1200 gpurun_out/${TAG}_bench_cfg3_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n$N.json"))
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "parity", d.get("parity_check"))
print(d["phases_ms"]); print(d["kernels_ms_per_step"])
PY
