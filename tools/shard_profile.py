"""One GPU: times the kernels of ONE key range of an N-way suffix-range sharding (what each rank of
an N-GPU build runs), device-resident input.
    python tools/shard_profile.py --nparts 8 --part 3 [--workload cfg3 --scale 1.0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bwtb3m_b200 import Engine, multigpu, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--nparts", type=int, default=8)
ap.add_argument("--part", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
itype, data, nsym = workloads.make(a.workload, a.scale)
dev = torch.from_numpy(data).cuda()
eng = Engine(0)
eng.load_device(dev.data_ptr(), dev.numel(), itype)
buf = multigpu.ShardBuffers(eng, 0, 32, 262144, False)
for k in range(a.steps + 1):
    if k == 1:
        eng.set_profile(True)
    eng.load_device(dev.data_ptr(), dev.numel(), itype)
    un = eng.shard_build(a.part, a.nparts, *buf.ptrs(), preisarate=buf.prerate)
    eng.sync()
kt = eng.kernel_times()
i = eng.info()
print("unresolved", un, "ms_sort %.3f" % i["ms_sort"])
for name, r in sorted(kt.items()):
    print("%-26s %3d launches  %8.3f ms/step" % (name, r["launches"], r["ms"] / a.steps))
