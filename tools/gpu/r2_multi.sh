# usage: gpurun --gpus N -- 'bash tools/gpu/r2_multi.sh N TAG [quick]'
# the multi-GPU build behind the C ABI (b3m_multi_*, bwtb3m ngpus=) and the process-per-GPU build (torch.distributed):
# parity tests on all N GPUs, then the N-GPU bench line (with parity_check and the e2e check)
N=${1:-2}; TAG=${2:-r2m}; QUICK=${3:-0}
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi_abi.py -x -q --tb=short 2>&1 | tail -12 | cut -c1-1200
if [ "$QUICK" = "0" ]; then
timeout 1500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_xshard.py -x -q --tb=short 2>&1 | tail -12 | cut -c1-1200
else
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q --tb=short -k "io or cfg3" 2>&1 | tail -12 | cut -c1-1200
fi
for G in $N; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $G --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_cfg3_n$G.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n$G.json
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_cfg3_n$G.err | tail -c 1500
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n$G.json"))
print("N=$G ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["e2e"].get("check"), "parity", d.get("parity_check"))
print(d["phases_ms"]); print(d["kernels_ms_per_step"])
PY
done
# the C++ host path end to end: file in, files out, N GPUs inside one process
python - <<PY
import os, subprocess, time, numpy as np, sys
sys.path.insert(0, ".")
from bwtb3m_b200 import workloads, MultiEngine
itype, data, nsym = workloads.make("cfg3", 1.0)
import torch
host = torch.from_numpy(data).pin_memory()
m = MultiEngine($N)
for k in range(4):
    t0 = time.perf_counter(); m.load_host_ptr(host.data_ptr(), host.numel(), itype); t1 = time.perf_counter(); m.build(); t2 = time.perf_counter()
    print("MultiEngine($N) cfg3: load %.2f ms build %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1)), m.stats())
m.close()
PY
