"""Runs the hot path a few times on device-resident input, nothing else: the command ncu wraps
(launch list with gpu__time_duration, and --set full captures of single kernels)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bwtb3m_b200 import Engine, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--numblocks", type=int, default=1)
ap.add_argument("--bwtonly", type=int, default=0)
a = ap.parse_args()
itype, data, nsym = workloads.make(a.workload, a.scale)
dev = torch.from_numpy(data).cuda()
eng = Engine(0)
for k in range(a.steps):
    eng.load_device(dev.data_ptr(), dev.numel(), itype)
    eng.build(numblocks=a.numblocks, bwtonly=bool(a.bwtonly))
    eng.sync()
i = eng.info()
print("n=%d launches=%d ms_total=%.3f sort=%.3f gap=%.3f merge=%.3f walk=%.3f" % (i["n"], i["launches"], i["ms_total"], i["ms_sort"], i["ms_gap"], i["ms_merge"], i["ms_walk"]))
