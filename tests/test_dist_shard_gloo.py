"""CPU, gloo, world size 2, 3 and 4: the suffix-range sharding driver (bwtb3m_b200.multigpu.build_sharded)
with a model engine that writes its slice from the oracle's suffix array -- exercises the zeroed
global-place buffers, the unresolved vote and the sum-reduce to rank 0."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ModelEngine:
    """shard_build / shard_finish on host memory: part p owns the ranks [p*n/P, (p+1)*n/P)."""

    def __init__(self, text, sa, bwt, isa, force_unresolved_on=None):
        self.t, self.sa, self.bwt, self.isa = text, sa, bwt, isa
        self.force = force_unresolved_on
        self.result = None
        self.stream_ptr = 0

    def info(self):
        return {"n": int(self.t.size)}

    def default_preisarate(self, bwtonly=False):
        return 64

    def shard_rows(self, nparts):
        n = int(self.t.size)
        return [p * n // nparts for p in range(nparts + 1)]

    @staticmethod
    def _view(ptr, n, dt):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dt)), shape=(n,))

    def shard_build(self, part, nparts, bwt_p, pre_p, sa_p, isa_p, spec_p, preisarate=0, sasamplingrate=32, isasamplingrate=262144, bwtonly=False):
        n = self.t.size
        lo, hi = part * n // nparts, (part + 1) * n // nparts
        self.rates = (preisarate, sasamplingrate, isasamplingrate)
        self._view(bwt_p, n, C.c_uint8)[lo:hi] = self.bwt[lo:hi]
        pre = self._view(pre_p, -(-n // preisarate), C.c_int32)
        sas = self._view(sa_p, -(-n // sasamplingrate), C.c_int64)
        isas = self._view(isa_p, -(-n // isasamplingrate), C.c_int64)
        for r in range(lo, hi):
            p = int(self.sa[r])
            if p % preisarate == 0:
                pre[p // preisarate] = r
            if p % isasamplingrate == 0:
                isas[p // isasamplingrate] = r
            if r % sasamplingrate == 0:
                sas[r // sasamplingrate] = p
        return 5 if self.force == part else 0

    def shard_finish(self, nparts, bwt_p, pre_p, sa_p, isa_p, spec_p):
        n = self.t.size
        pr, sr, ir = self.rates
        self.result = (self._view(bwt_p, n, C.c_uint8).copy(), self._view(pre_p, -(-n // pr), C.c_int32).copy(),
                       self._view(sa_p, -(-n // sr), C.c_int64).copy(), self._view(isa_p, -(-n // ir), C.c_int64).copy())


def _worker(rank, world, port, force, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bwtb3m_b200 import multigpu
    from oracle import oracle as orc
    orc.build()
    t = np.random.default_rng(3).integers(0, 4, size=5003, dtype=np.uint8)
    sa = orc.sa_circular(t)
    bwt, isa = orc.bwt_from_sa(t, sa)
    eng = ModelEngine(t, sa, bwt, isa, force)
    buf = multigpu.ShardBuffers(eng, 64, 8, 32, False, device=torch.device("cpu"))
    ok, _ = multigpu.build_sharded(eng, 64, 8, 32, False, buffers=buf, rank=rank, world=world)
    if rank == 0:
        if ok:
            b, pre, s, i = eng.result
            good = (np.array_equal(b, bwt) and np.array_equal(pre, isa[::64].astype(np.int32)) and
                    np.array_equal(s, sa[::8].astype(np.int64)) and np.array_equal(i, isa[::32].astype(np.int64)))
            q.put(("ok", bool(good)))
        else:
            q.put(("unresolved", eng.result is None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,force", [(2, None), (3, None), (4, None), (2, 1)])
def test_sharded_driver_gloo(world, force):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 7 + (force or 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, force, q)) for r in range(world)]
    for p in procs:
        p.start()
    kind, good = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert good
    assert kind == ("unresolved" if force is not None else "ok")
