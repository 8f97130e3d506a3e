"""CPU, world_size 2..4 over gloo: the multi-GPU schedule of bwtb3m_b200/multigpu.py (who
broadcasts, reduces, sends what; chain split; gt-bit bookkeeping; anchor maps) driven with a
small numpy model of the per-rank primitives, checked against the CPU oracle.  The CUDA engine
plugs into the same driver through EngineOps (tests/test_gpu_dist.py, tools/dist_check.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyOps:
    """Model of EngineOps on a circular text without terminator."""

    def __init__(self, text, prerate):
        self.t = np.asarray(text, dtype=np.uint8)
        self.n = int(self.t.size)
        self.prerate = prerate
        self.npre = (self.n + prerate - 1) // prerate
        self.leaves = []
        tt = self.t.tobytes()
        self.rot = lambda i: tt[i:] + tt[:i]

    def zeros(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    def begin(self, gt, prerank, rsamp):
        self.gt, self.prerank, self.rsamp = gt.numpy(), prerank.numpy(), rsamp.numpy()

    def build_range(self, a0, a1, nblocks, L):
        sa = sorted(range(a0, a1), key=self.rot)
        rank = {p: k for k, p in enumerate(sa)}
        Ln = L.numpy()
        for k, p in enumerate(sa):
            Ln[k] = 0 if p == a0 else self.t[p - 1]
        for p in range(a0, a1):
            self.gt[p] = rank[p] > rank[a0]
            if p % self.prerate == 0:
                self.prerank[p // self.prerate] = rank[p]
        self.leaves.append((a0, a1, sa))
        return rank[a0]

    def chains(self, nr):
        nch = min(max(nr // 3, 1), 7)
        chl = -(-nr // nch)
        return chl, -(-nr // chl)

    def zranks(self, a0, a1, r1, chl, nch, r0):
        r0n = r0.numpy()
        for c in range(nch):
            z = min(a1 + (c + 1) * chl, r1) % self.n
            rz = self.rot(z)
            for (s, e, sa) in self.leaves:
                if a0 <= s < a1:
                    r0n[c] += sum(1 for p in sa if p != z and self.rot(p) < rz)

    def gap(self, LA, a0, na, termA, r1, chl, nch, c_lo, c_hi, r0, gtnew, G):
        t, n = self.t, self.n
        a1 = a0 + na
        La = LA.numpy()[:na]
        Gn, gn, r0n = G.numpy(), gtnew.numpy(), r0.numpy()
        sigma = int(t.max()) + 1
        cnt = np.bincount(t[a0:a1], minlength=sigma)
        CA = np.concatenate([[0], np.cumsum(cnt)])
        lastA = t[a1 - 1]
        for c in range(c_lo, min(c_hi, nch)):
            zlo, zhi = a1 + c * chl, min(a1 + (c + 1) * chl, r1)
            r = int(r0n[c])
            for j in range(zhi, zlo, -1):
                gtj = self.gt[j] if j < r1 else (self.rot(r1 % n) > self.rot(a1))
                c0 = t[j - 1]
                occ = int(np.count_nonzero(La[:r] == c0)) - (1 if (termA < r and c0 == 0) else 0)
                r = int(CA[c0]) + occ + (1 if (c0 == lastA and gtj) else 0)
                Gn[r] += 1
                gn[j - 1 - a1] = r > termA
                if (j - 1) % self.prerate == 0:
                    self.rsamp[(j - 1) // self.prerate] = r

    def merge(self, LA, na, termA, LR, nr, termR, a1, G, LM):
        La, Lr, Lm, Gn = LA.numpy(), LR.numpy(), LM.numpy(), G.numpy()
        Lr[termR] = self.t[a1 - 1]
        S = np.cumsum(Gn[: na + 1])
        assert S[-1] == nr
        for k in range(na + 1):
            q = S[k] - Gn[k]
            o = k + q
            Lm[o:o + Gn[k]] = Lr[q:q + Gn[k]]
            if k < na:
                Lm[o + Gn[k]] = La[k]
        Gn[: na + 1] = S
        return int(termA + S[termA])

    def merge_samples(self, a0, a1, r1, G):
        Gn = G.numpy()
        for q in range(-(-a0 // self.prerate), -(-r1 // self.prerate)):
            k = self.prerank[q]
            self.prerank[q] = k + (Gn[k] if q * self.prerate < a1 else self.rsamp[q])

    def finish(self, L, term, q_lo, q_hi, sarate, isarate, bwtonly, numblocks):
        n = self.n
        bw = L.numpy()[:n].copy()
        bw[term] = self.t[n - 1]
        self.bwt = bw
        if bwtonly:
            return None, None
        sigma = int(bw.max()) + 1
        Cc = np.concatenate([[0], np.cumsum(np.bincount(bw, minlength=sigma))])
        sa = np.full((n + sarate - 1) // sarate, -1, dtype=np.int64)
        isa = np.full((n + isarate - 1) // isarate, -1, dtype=np.int64)
        for q in range(q_lo, min(q_hi, self.npre)):
            p, r = q * self.prerate, int(self.prerank[q])
            steps = self.prerate if q else n - (self.npre - 1) * self.prerate
            for _ in range(steps):
                if p % isarate == 0:
                    isa[p // isarate] = r
                if r % sarate == 0:
                    sa[r // sarate] = p
                p = p - 1 if p else n - 1
                r = int(Cc[bw[r]]) + int(np.count_nonzero(bw[:r] == bw[r]))
        return torch.from_numpy(sa), torch.from_numpy(isa)


def _worker(rank, world, port, text, prerate, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import sys
        sys.path.insert(0, ROOT)
        from bwtb3m_b200.multigpu import DistBuild
        ops = NumpyOps(text, prerate)
        drv = DistBuild(ops)
        res = drv.build(1, sasamplingrate=4, isasamplingrate=8, bwtonly=False)
        if rank == 0:
            q.put(("ok", ops.bwt, ops.prerank.copy(), res["sa"].numpy().copy(), res["isa"].numpy().copy(), drv.stats["merges"]))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put(("err", rank, traceback.format_exc()))
        raise


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(world, text, prerate):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, text, prerate, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert out[0] == "ok", out
    return out[1:]


@pytest.mark.parametrize("world,n,sigma,seed", [(2, 301, 4, 1), (2, 64, 2, 2), (3, 500, 4, 3), (4, 777, 5, 4)])
def test_distributed_schedule_gloo(oracle, world, n, sigma, seed):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, sigma, size=n, dtype=np.uint8)
    t[-1] = sigma  # make the text primitive
    prerate = 16
    bwt, prerank, sa_s, isa_s, merges = run_world(world, t, prerate)
    sa = oracle.sa_circular(t)
    rb, isa = oracle.bwt_from_sa(t, sa)
    assert np.array_equal(bwt, rb)
    assert np.array_equal(prerank.astype(np.int64), isa[::prerate].astype(np.int64))
    assert np.array_equal(sa_s, sa[::4].astype(np.int64))
    assert np.array_equal(isa_s, isa[::8].astype(np.int64))
    assert merges >= 1


def test_tree_groups_and_ranges():
    import sys
    sys.path.insert(0, ROOT)
    from bwtb3m_b200.multigpu import block_range, tree_groups
    assert tree_groups(1) == []
    assert tree_groups(2) == [(0, 2)]
    assert tree_groups(8) == [(0, 8), (0, 4), (0, 2), (2, 4), (4, 8), (4, 6), (6, 8)]
    assert tree_groups(3) == [(0, 3), (1, 3)]
    n = 1001
    cover = [block_range(n, 8, i) for i in range(8)]
    assert cover[0][0] == 0 and cover[-1][1] == n
    assert all(cover[i][1] == cover[i + 1][0] for i in range(7))
