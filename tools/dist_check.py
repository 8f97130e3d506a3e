"""torchrun script: multi-GPU build (bwtb3m_b200.multigpu, NCCL) == single-GPU build, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [--workload cfg2 --scale 0.1 | --nsym 200003 --itype pacterm] [--local-blocks 1]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bwtb3m_b200 import Engine, multigpu, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--nsym", type=int, default=200_003)
ap.add_argument("--itype", default="pacterm")
ap.add_argument("--local-blocks", type=int, default=1)
ap.add_argument("--seed", type=int, default=7)
ap.add_argument("--strategy", default="auto", choices=["auto", "shard", "merge"])
ap.add_argument("--io", action="store_true", help="input and results over every rank's PCIe link (multigpu.load_distributed / fetch_distributed; pacterm, sharded)")
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if a.workload:
    itype, data, nsym = workloads.make(a.workload, a.scale)
elif a.itype in ("pac", "pacterm"):
    itype, data = a.itype, workloads.random_pac(a.nsym, a.seed)
else:
    itype, data = "bytestream", np.random.default_rng(a.seed).integers(0, 256, size=a.nsym, dtype=np.uint8)

stream = torch.cuda.Stream()
eng = Engine(local, stream.cuda_stream)
io_state, shared = {}, None
host = torch.from_numpy(data).pin_memory()
if not a.io:
    eng.load_host(data, itype)
drv = None
for it in range(2):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        if a.io:
            multigpu.load_distributed(eng, host, itype, io_state)
            if shared is not None and isinstance(drv, dict):
                drv["stream_sa_host"] = shared["sa"].ptr()  # second build: the SA samples leave during the finish kernels
                torch.cuda.synchronize()
                if rank == 0:
                    shared["sa"].t.fill_(0xEE)  # what the first round's fetch left must not pass for this round's delivery
                    shared["bwa"].t.fill_(0xEE)
                dist.barrier()
        drv, res = multigpu.build_distributed(eng, local_blocks=a.local_blocks, sasamplingrate=32, isasamplingrate=1024, driver=drv, strategy=a.strategy)
        if a.io:
            i0 = eng.info()
            if shared is None:
                nw = (i0["n"] - 1 + 15) >> 4
                shared = {"bwa": multigpu.SharedHost(4 * nw, "bwa", rank, world, width=4), "sa": multigpu.SharedHost(8 * i0["nsa"], "sa", rank, world, width=8)}
                shared["bwa"].t.fill_(0xEE)
                shared["sa"].t.fill_(0xEE)
            multigpu.fetch_distributed(eng, drv, shared["bwa"].ptr(), shared["sa"].ptr())
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
ok = True
if rank == 0:
    multi = eng.fetch()
    info = eng.info()
    ref = Engine(local)
    ref.load_host(data, itype)
    ref.build(numblocks=1, sasamplingrate=32, isasamplingrate=1024, preisarate=info["preisarate"])
    one = ref.fetch()
    for k in ("bwt", "preisa", "sa", "isa"):
        same = np.array_equal(multi[k], one[k])
        ok = ok and same
        print("%s: %s" % (k, "equal" if same else "DIFFERENT"))
    if a.io:
        words, primary, l2, seq_len = ref.fetch_bwa()
        same = np.array_equal(shared["bwa"].t.numpy().view(np.uint32)[:words.size], words)
        ok = ok and same
        print("shared host BWA words: %s" % ("equal" if same else "DIFFERENT"))
        same = np.array_equal(shared["sa"].t.numpy().view(np.uint64)[:one["sa"].size], one["sa"])
        ok = ok and same
        print("shared host SA samples: %s" % ("equal" if same else "DIFFERENT"))
    print("strategy used: %s" % res["strategy"])
    print("world=%d n=%d second build %.1f ms; phases ms sort=%.1f gap=%.1f merge=%.1f walk=%.1f  %s" %
          (world, info["n"], dt * 1e3, info["ms_sort"], info["ms_gap"], info["ms_merge"], info["ms_walk"], "DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
