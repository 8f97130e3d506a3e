# usage: gpurun --gpus N -- 'bash tools/gpu/r2_n8b.sh N TAG': the N-GPU bench line only (per-rank kernel maxima, build timeline, e2e phases)
N=${1:-8}; TAG=${2:-r2q}
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_cfg3_n$N.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n$N.json
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_cfg3_n$N.err | tail -c 1200
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n$N.json"))
print("N=$N ms/step", d["ms_per_step"], "e2e", d["e2e"], "parity", d.get("parity_check",{}).get("ok"))
print(d["phases_ms"]); print(d["kernels_ms_per_step"]); print(d.get("kernels_ms_per_step_max_over_ranks")); print(d.get("build_timeline_ms"))
PY
