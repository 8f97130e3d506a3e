"""CPU, gloo, world size 2, 3 and 4: the position-sharded driver (bwtb3m_b200.multigpu.build_xsharded, the default
multi-GPU path for texts over at most four codes) with a model engine on POSIX shared memory:
count own text positions -> all-gather of the bin totals -> scatter, every rank storing records straight into the
record array of the rank that owns their key range -> all-reduce fence -> every rank orders its key range and stores
its BWT rows / anchors / samples into rank 0's buffers -> all-reduce vote -> rank 0 adopts.  Also the vote failing,
and the "samples were sent to the host during the finish" flag travelling with the vote (all ranks or none)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

NB = 16  # level-1 bins of the model: the first two symbols of a suffix


class XModelEngine:
    """xshard_count / xshard_scatter / xshard_finish of bwtb3m_b200.engine.Engine on host memory.  A record is one
    uint64: (bin << 32) | suffix index; inside a key range the records lie bin after bin, inside a bin rank after rank
    (the layout the real scatter kernel produces from the all-gathered totals)."""

    def __init__(self, text, sa, bwt, isa, force_unresolved_on=None, no_stream_on=None):
        self.t, self.sa, self.bwt, self.isa = text, sa, bwt, isa
        self.force, self.no_stream_on = force_unresolved_on, no_stream_on
        self.result = None
        self.stream_ptr = 0
        self.armed = None
        self.delivered = False
        self.streamed_to = None

    def info(self):
        return {"n": int(self.t.size)}

    def default_preisarate(self, bwtonly=False):
        return 64

    @staticmethod
    def _view(ptr, n, dt):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dt)), shape=(n,))

    def _bins(self, lo, hi):
        n = self.t.size
        i = np.arange(lo, hi)
        return (self.t[i].astype(np.int64) << 2) | self.t[(i + 1) % n]

    def xshard_count(self, part, nparts, totals_ptr, preisarate=0, sasamplingrate=32, isasamplingrate=262144, bwtonly=False):
        n = self.t.size
        self.part, self.nparts = part, nparts
        self.rates = (preisarate, sasamplingrate, isasamplingrate)
        self.lo, self.hi = part * n // nparts, (part + 1) * n // nparts
        self.mybins = self._bins(self.lo, self.hi)
        tot = self._view(totals_ptr, 2048, C.c_int64)
        tot[:NB] = np.bincount(self.mybins, minlength=NB)
        return NB

    def xshard_scatter(self, h, ptrs, caps):
        h = np.asarray(h, dtype=np.int64)  # [world, NB]
        n, P = self.t.size, self.nparts
        total = h.sum(axis=0)
        # bins cut into P key ranges of about n / P suffixes (the rule of the real engine)
        bnd, first, acc, p = [0], [0], 0, 1
        for b in range(NB):
            while p < P and acc >= n * p // P:
                bnd.append(b); first.append(acc); p += 1
            acc += int(total[b])
        while len(bnd) < P + 1:
            bnd.append(NB); first.append(n)
        bnd[P], first[P] = NB, n
        self.bnd, self.first = bnd, first
        for q in range(P):
            if first[q + 1] - first[q] > caps[q]:
                raise RuntimeError("a key range exceeds its record array")
        order = np.argsort(self.mybins, kind="stable")
        for q in range(P):
            out = self._view(ptrs[q], caps[q], C.c_uint64)
            base = 0
            for b in range(bnd[q], bnd[q + 1]):
                mine = order[self.mybins[order] == b]
                at = base + int(h[:self.part, b].sum())
                out[at:at + mine.size] = (np.uint64(b) << np.uint64(32)) | (mine + self.lo).astype(np.uint64)
                base += int(total[b])

    def xshard_stream_sa(self, d_local, host):
        self.armed = (d_local, host)

    def xshard_sa_delivered(self):
        return self.delivered

    def xshard_finish(self, own_ptr, bwt_p, pre_p, sa_p, isa_p, spec_p):
        n = self.t.size
        pr, sr, ir = self.rates
        m = self.first[self.part + 1] - self.first[self.part]
        recs = self._view(own_ptr, max(m, 1), C.c_uint64)[:m]
        idx = (recs & np.uint64(0xffffffff)).astype(np.int64)
        ranks = np.sort(self.isa[idx].astype(np.int64))  # the key range is a contiguous range of ranks
        assert m == 0 or (ranks[0] == self.first[self.part] and ranks[-1] == self.first[self.part + 1] - 1 and np.all(np.diff(ranks) == 1))
        b = self._view(bwt_p, n, C.c_uint8)
        pre = self._view(pre_p, -(-n // pr), C.c_int32)
        sas = self._view(sa_p, -(-n // sr), C.c_int64)
        isas = self._view(isa_p, -(-n // ir), C.c_int64)
        armed, self.armed = self.armed, None
        self.delivered = armed is not None and self.no_stream_on != self.part
        host = self._view(armed[1], -(-n // sr), C.c_int64) if self.delivered else None
        for r in ranks:
            p = int(self.sa[r])
            b[r] = self.bwt[r]
            if p % pr == 0:
                pre[p // pr] = r
            if p % ir == 0:
                isas[p // ir] = r
            if r % sr == 0:
                sas[r // sr] = p
                if host is not None:
                    host[r // sr] = p
        return 7 if self.force == self.part else 0

    # ---- what load_distributed / fetch_distributed call ----
    def load_device(self, ptr, n, inputtype):
        self.loaded = self._view(ptr, n, C.c_uint8).copy()

    def pack_bwa(self, words_ptr, w_lo, w_hi):
        self._view(words_ptr, len(self.words), C.c_uint32)[w_lo:w_hi] = self.words[w_lo:w_hi]

    def fetch_ptrs(self, bwt_p, pre_p, sa_p, isa_p):
        self.small_fetch = (pre_p, isa_p)

    def shard_adopt(self, nparts, bwt_p, pre_p, sa_p, isa_p, spec_p):
        n = self.t.size
        pr, sr, ir = self.rates
        self.result = (self._view(bwt_p, n, C.c_uint8).copy(), self._view(pre_p, -(-n // pr), C.c_int32).copy(),
                       self._view(sa_p, -(-n // sr), C.c_int64).copy(), self._view(isa_p, -(-n // ir), C.c_int64).copy())


def _worker(rank, world, port, force, no_stream_on, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bwtb3m_b200 import multigpu
    from oracle import oracle as orc
    from test_dist_direct_gloo import ShmMemory
    orc.build()
    t = np.random.default_rng(6).integers(0, 4, size=9001, dtype=np.uint8)
    sa = orc.sa_circular(t)
    bwt, isa = orc.bwt_from_sa(t, sa)
    eng = XModelEngine(t, sa, bwt, isa, force, no_stream_on)
    mem = ShmMemory()
    res = multigpu.DirectResults(eng, 64, 32, 32, False, rank, world, mem=mem)
    xr = multigpu.XRecs(eng, rank, world, mem=mem)
    host = res.extra("host_sa_model", 8 * (-(-t.size // 32)))  # stands in for the SharedHost buffer of the samples
    good = True
    for it in range(3):  # buffers are reused; the samples are "sent to the host" from the second build on
        res.stream_sa_host = host if it else 0
        ok = multigpu.build_xsharded(eng, res, xr, 32, 32, False, device=torch.device("cpu"))
        if force is not None:
            good = good and ok is False and not res.sa_streamed
        else:
            want_streamed = bool(it) and no_stream_on is None
            good = good and ok is True and res.sa_streamed == want_streamed
            if rank == 0:
                b, pre, s, i = eng.result
                good = good and (np.array_equal(b, bwt) and np.array_equal(pre, isa[::64].astype(np.int32)) and
                                 np.array_equal(s, sa[::32].astype(np.int64)) and np.array_equal(i, isa[::32].astype(np.int64)))
                if want_streamed:
                    hs = XModelEngine._view(host, -(-t.size // 32), C.c_int64)
                    good = good and np.array_equal(hs, sa[::32].astype(np.int64))
        dist.barrier()
    # ---- the input over every rank's link, the two large results back the same way (CPU stand-ins for the copies) ----
    if force is None:
        src = np.random.default_rng(9).integers(0, 256, size=10_007, dtype=np.uint8)
        io = {}
        sent = multigpu.load_distributed(eng, torch.from_numpy(src), "pacterm", io, device=torch.device("cpu"))
        good = good and np.array_equal(eng.loaded, src) and 0 < sent <= -(-src.size // world) + 256
        nw = (t.size - 1 + 15) >> 4
        eng.words = np.random.default_rng(10).integers(0, 1 << 32, size=nw, dtype=np.uint64).astype(np.uint32)
        state = {"direct": res}
        hw = res.extra("host_words_model", 4 * nw)
        hs = res.extra("host_sa_model2", 8 * (-(-t.size // 32)))
        for streamed in (False, True):
            XModelEngine._view(hw, nw, C.c_uint32)[:] = 0
            XModelEngine._view(hs, -(-t.size // 32), C.c_int64)[:] = -1
            dist.barrier()
            res.sa_streamed, res.stream_sa_host = streamed, hs
            multigpu.fetch_distributed(eng, state, hw, hs, 1, 1, device=torch.device("cpu"))
            dist.barrier()
            if rank == 0:
                good = good and np.array_equal(XModelEngine._view(hw, nw, C.c_uint32), eng.words) and eng.small_fetch == (1, 1)
                got = XModelEngine._view(hs, -(-t.size // 32), C.c_int64)
                if streamed:  # only sample 0 is written by the fetch, the rest is the finish kernels' business
                    good = good and got[0] == sa[0] and np.all(got[1:] == -1)
                else:
                    good = good and np.array_equal(got, sa[::32].astype(np.int64))
            dist.barrier()  # the other ranks clear the buffers for the next round only after rank 0 has looked
    flags = [None] * world
    dist.all_gather_object(flags, bool(good))
    if rank == 0:
        q.put(all(flags))
    dist.barrier()
    xr.close()
    res.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,force,no_stream_on", [(2, None, None), (3, None, None), (4, None, None), (3, 1, None), (3, None, 2)])
def test_xsharded_driver_gloo(world, force, no_stream_on):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + world * 11 + (force or 0) + 3 * (no_stream_on or 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, force, no_stream_on, q)) for r in range(world)]
    for p in procs:
        p.start()
    good = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert good
