"""CPU, gloo, world size 2, 3 and 4: the direct-store sharding driver (bwtb3m_b200.multigpu.DirectResults +
build_sharded_direct).  POSIX shared memory stands in for rank 0's HBM mapped through CUDA IPC: every rank's
model engine writes its slice straight into rank 0's buffers, the only collective is the all-reduce of the
"unresolved" counts, rank 0 adopts the buffers."""
import ctypes as C
import os
import sys
from multiprocessing import shared_memory

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


class ShmMemory:
    """alloc / export / open / free / close of bwtb3m_b200.engine.DeviceMemory on shared host memory."""

    def __init__(self):
        self.seg = {}

    def _addr(self, shm):
        return C.addressof(C.c_char.from_buffer(shm.buf))

    def alloc(self, nbytes):
        shm = shared_memory.SharedMemory(create=True, size=max(nbytes, 1))
        shm.buf[:] = b"\xee" * len(shm.buf)  # the driver must not rely on zeroed buffers
        a = self._addr(shm)
        self.seg[a] = shm
        return a

    def export(self, ptr):
        return self.seg[ptr].name.encode()

    def open(self, handle):
        shm = shared_memory.SharedMemory(name=handle.decode())
        a = self._addr(shm)
        self.seg[a] = shm
        return a

    def close(self, ptr):
        self.seg.pop(ptr).close()

    def copy(self, dst, src, nbytes, stream_ptr=0):
        C.memmove(dst, src, nbytes)

    def free(self, ptr):
        shm = self.seg.pop(ptr)
        shm.close()
        shm.unlink()


def _worker(rank, world, port, force, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bwtb3m_b200 import multigpu
    from oracle import oracle as orc
    from test_dist_shard_gloo import ModelEngine
    orc.build()
    t = np.random.default_rng(4).integers(0, 4, size=7001, dtype=np.uint8)
    sa = orc.sa_circular(t)
    bwt, isa = orc.bwt_from_sa(t, sa)
    eng = ModelEngine(t, sa, bwt, isa, force)
    eng.shard_adopt = eng.shard_finish
    res = multigpu.DirectResults(eng, 64, 8, 32, False, rank, world, mem=ShmMemory())
    good = True
    for it in range(2):  # the buffers are reused by the next build
        ok = multigpu.build_sharded_direct(eng, res, 8, 32, False, device=torch.device("cpu"))
        if rank == 0:
            if force is not None:
                good = good and not ok
            else:
                b, pre, s, i = eng.result
                good = good and ok and (np.array_equal(b, bwt) and np.array_equal(pre, isa[::64].astype(np.int32)) and
                                        np.array_equal(s, sa[::8].astype(np.int64)) and np.array_equal(i, isa[::32].astype(np.int64)))
        dist.barrier()
    if rank == 0:
        q.put(bool(good))
    dist.barrier()
    res.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,force", [(2, None), (3, None), (4, None), (3, 2)])
def test_direct_sharded_driver_gloo(world, force):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 7 + (force or 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, force, q)) for r in range(world)]
    for p in procs:
        p.start()
    good = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert good


def test_rank_slices_cover_the_results_exactly_once():
    """fetch_distributed / SharedHost: the slices the ranks send to the host partition [0, count), start on multiples of
    64 elements, and the slice length of rank 0 bounds every other one (it sizes the staging buffers)."""
    from bwtb3m_b200 import multigpu
    for count in (0, 1, 63, 64, 65, 6251, 1 << 20, 193_750_001):
        for world in (1, 2, 3, 4, 8):
            cover = 0
            per = multigpu.rank_slice(count, 0, world)[1]
            for r in range(world):
                lo, hi = multigpu.rank_slice(count, r, world)
                assert lo == min(cover, count) and lo % 64 == 0 or lo == count
                assert hi - lo <= max(per, 0)
                cover = max(cover, hi)
            assert cover == count
