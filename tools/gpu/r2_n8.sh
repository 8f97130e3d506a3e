# usage: gpurun --gpus 8 -- 'bash tools/gpu/r2_n8.sh 8 TAG'   (kept short: an 8-GPU minute costs eight)
N=${1:-8}; TAG=${2:-r2o}
set -x
timeout 600 python -m pytest tests/test_gpu_multi_abi.py -x -q --tb=short 2>&1 | tail -8 | cut -c1-1200
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q --tb=short -k "100003" 2>&1 | tail -8 | cut -c1-1200
for WL in cfg3 cfg5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --workload $WL --gpus $N --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_${WL}_n$N.err | grep "^{" > gpurun_out/${TAG}_bench_${WL}_n$N.json
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_${WL}_n$N.err | tail -c 1200
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${WL}_n$N.json"))
print("$WL N=$N ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["e2e"].get("check"), "parity", d.get("parity_check"))
print(d["phases_ms"]); print(d["kernels_ms_per_step"])
PY
done
python - <<PY
import time, sys
sys.path.insert(0, ".")
from bwtb3m_b200 import workloads, MultiEngine
itype, data, nsym = workloads.make("cfg3", 1.0)
import torch
host = torch.from_numpy(data).pin_memory()
m = MultiEngine($N)
for k in range(3):
    t0 = time.perf_counter(); m.load_host_ptr(host.data_ptr(), host.numel(), itype); t1 = time.perf_counter(); m.build(); t2 = time.perf_counter()
    print("MultiEngine($N) cfg3: load %.2f ms build %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1)), m.stats())
m.close()
PY
