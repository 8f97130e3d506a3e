"""Multi-GPU driver: one process per GPU, text replicated, rank i sorts text range i, the balanced
merge tree runs over torch.distributed (NCCL on GPUs, gloo in the CPU tests of the schedule).

Replaces the block scheduling + merge tree of BwtMergeSortTemplate::computeBwt (libmaus2, reached
from /root/reference/src/bwtb3m.cpp:63) for the case "blocks sharded over the GPUs of one box"
(SURVEY.md 8e).  Per merge of A=[a0,a1) (ranks lo..mid-1, leader lo) and R=[a1,r1) (ranks
mid..hi-1, leader mid):

  1. broadcast L_A from lo and gt[a1:r1] from mid to the whole group,
  2. every rank of A's side adds the z-ranks of its own leaves, all-reduce(sum) -> start ranks,
  3. the chains of R's text are split evenly over ALL ranks of the group (K5 on each),
  4. reduce(sum) the partial gap arrays and anchor ranks to lo, send L_R and R's anchors to lo,
  5. lo merges (K6); if the node is needed as a right part later, its new gt bits are
     all-reduced over the group.

The arithmetic is behind an `ops` object: EngineOps (CUDA engine through the C ABI) on GPUs; the
CPU tests plug in a numpy model of the same primitives to exercise this schedule under gloo.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist


def block_range(n, world, i):
    bs = (n + world - 1) // world
    return min(i * bs, n), min((i + 1) * bs, n)


def tree_groups(world):
    """All rank intervals [lo,hi) of the balanced merge tree with more than one rank, in a fixed
    order (every rank must create the process groups in the same order)."""
    out = []

    def rec(lo, hi):
        if hi - lo <= 1:
            return
        out.append((lo, hi))
        mid = (lo + hi) // 2
        rec(lo, mid)
        rec(mid, hi)

    rec(0, world)
    return out


class DistBuild:
    """Runs the distributed build on the calling rank.  ops: see EngineOps."""

    def __init__(self, ops, rank=None, world=None):
        self.ops = ops
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.groups, self.pairs = {}, {}
        for lo, hi in tree_groups(self.world):
            g = dist.new_group(ranks=list(range(lo, hi)))
            self.groups[(lo, hi)] = g
            # the two leaders of a merge exchange L_R and R's anchors inside their own 2-rank group
            self.pairs[(lo, hi)] = g if hi - lo == 2 else dist.new_group(ranks=[lo, (lo + hi) // 2])
        self.stats = {"bytes_sent": 0, "merges": 0}

    # -- 32-bit unsigned counters travel as int32 tensors: NCCL has no uint32, and a two's
    #    complement sum has the same bits --
    @staticmethod
    def _i32(t):
        return t

    def _bcast(self, t, src, group):
        dist.broadcast(self._i32(t), src=src, group=group)

    def _allreduce_sum(self, t, group):
        dist.all_reduce(self._i32(t), op=dist.ReduceOp.SUM, group=group)

    def _reduce_sum(self, t, dst, group):
        dist.reduce(self._i32(t), dst=dst, op=dist.ReduceOp.SUM, group=group)

    def build(self, local_blocks=1, sasamplingrate=32, isasamplingrate=262144, bwtonly=False):
        ops, r, W = self.ops, self.rank, self.world
        n = ops.n
        if n < 2 * W:
            raise ValueError("text too short for %d ranks" % W)
        self.gt = ops.zeros(n, torch.uint8)
        self.prerank = ops.zeros(ops.npre, torch.int32)
        self.rsamp = ops.zeros(ops.npre, torch.int32)
        ops.begin(self.gt, self.prerank, self.rsamp)
        node = self._rec(0, W, False, local_blocks)
        # root: every rank gets the final BWT and the anchors, walks its share of the chains
        L = node["L"] if r == 0 else ops.zeros(n + 16, torch.uint8)
        meta = ops.zeros(1, torch.int64)
        if r == 0:
            meta[0] = node["term"]
        if W > 1:
            dist.broadcast(L, src=0)
            dist.broadcast(meta, src=0)
            self._bcast(self.prerank, 0, None)
        term = int(meta[0].item())
        q_lo, q_hi = ops.npre * r // W, ops.npre * (r + 1) // W
        sa, isa = ops.finish(L, term, q_lo, q_hi, sasamplingrate, isasamplingrate, bwtonly, W * local_blocks)
        if not bwtonly and W > 1:
            # unset entries are ~0 = -1 as int64; every entry is set by exactly one rank
            dist.reduce(sa.view(torch.int64), dst=0, op=dist.ReduceOp.MAX)
            dist.reduce(isa.view(torch.int64), dst=0, op=dist.ReduceOp.MAX)
        return {"L": L, "term": term, "sa": sa, "isa": isa}

    def _rec(self, lo, hi, need_gt, local_blocks):
        ops, r, n, W = self.ops, self.rank, self.ops.n, self.world
        if hi - lo == 1:
            a0, a1 = block_range(n, W, lo)
            L = ops.zeros(a1 - a0 + 16, torch.uint8)
            term = ops.build_range(a0, a1, local_blocks, L)
            return {"a0": a0, "a1": a1, "L": L, "term": term}
        mid = (lo + hi) // 2
        mine = self._rec(lo, mid, need_gt, local_blocks) if r < mid else self._rec(mid, hi, True, local_blocks)
        group = self.groups[(lo, hi)]
        a0, a1, r1 = block_range(n, W, lo)[0], block_range(n, W, mid)[0], block_range(n, W, hi - 1)[1]
        na, nr = a1 - a0, r1 - a1
        # 1. placeholders rows, L_A and R's gt bits to everybody in the group
        meta = ops.zeros(2, torch.int64)
        if r == lo:
            meta[0] = mine["term"]
        if r == mid:
            meta[1] = mine["term"]
        dist.all_reduce(meta, op=dist.ReduceOp.SUM, group=group)
        termA, termR = int(meta[0].item()), int(meta[1].item())
        LA = mine["L"] if r == lo else ops.zeros(na + 16, torch.uint8)
        dist.broadcast(LA, src=lo, group=group)
        dist.broadcast(self.gt[a1:r1], src=mid, group=group)
        self.stats["bytes_sent"] += (na + 16 + nr) * (hi - lo - 1) if r in (lo, mid) else 0
        # 2. start ranks of the chains
        chl, nch = ops.chains(nr)
        r0 = ops.zeros(nch, torch.int32)
        if r < mid:
            ops.zranks(a0, a1, r1, chl, nch, r0)
        self._allreduce_sum(r0, group)
        # 3. K5 on this rank's share of the chains
        P, g = hi - lo, r - lo
        c_lo, c_hi = nch * g // P, nch * (g + 1) // P
        G = ops.zeros(na + 1, torch.int32)
        gtnew = ops.zeros(nr, torch.uint8)
        qA1, qR1 = -(-a1 // ops.prerate), -(-r1 // ops.prerate)
        self.rsamp[qA1:qR1].zero_()
        ops.gap(LA, a0, na, termA, r1, chl, nch, c_lo, c_hi, r0, gtnew, G)
        # 4. partial gap arrays and anchor ranks to the leader; L_R and R's anchors to the leader
        self._reduce_sum(G, lo, group)
        if qR1 > qA1:
            self._reduce_sum(self.rsamp[qA1:qR1], lo, group)
        out = None
        pair = self.pairs[(lo, hi)]
        if r == mid:
            dist.broadcast(mine["L"], src=mid, group=pair)
            if qR1 > qA1:
                dist.broadcast(self.prerank[qA1:qR1], src=mid, group=pair)
            self.stats["bytes_sent"] += nr + 16
        if r == lo:
            LR = ops.zeros(nr + 16, torch.uint8)
            dist.broadcast(LR, src=mid, group=pair)
            if qR1 > qA1:
                dist.broadcast(self.prerank[qA1:qR1], src=mid, group=pair)
            LM = ops.zeros(na + nr + 16, torch.uint8)
            termM = ops.merge(LA, na, termA, LR, nr, termR, a1, G, LM)
            ops.merge_samples(a0, a1, r1, G)
            out = {"a0": a0, "a1": r1, "L": LM, "term": termM}
            self.stats["merges"] += 1
        # 5. gt bits of the merged node, if it (or an ancestor reached through left links) is a right part later
        if need_gt:
            dist.all_reduce(gtnew, op=dist.ReduceOp.SUM, group=group)
            self.gt[a1:r1].copy_(gtnew)
        return out


class _DevArray:
    """Zero-copy view of engine-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class EngineOps:
    """The primitives of the build on one GPU, through the b3m_engine_blk_* C ABI."""

    def __init__(self, engine, preisarate=0, largelcpthres=16384):
        self.e = engine
        self.device = torch.device("cuda", torch.cuda.current_device())
        i = engine.info()
        self.n = i["n"]
        self._preisarate = preisarate
        self._largelcpthres = largelcpthres
        # the anchor spacing must be known before the buffers are allocated: same rule as the engine
        self.prerate = preisarate or engine.default_preisarate()
        self.npre = (self.n + self.prerate - 1) // self.prerate

    def zeros(self, n, dtype):
        return torch.zeros(n, dtype=dtype, device=self.device)

    def _chk(self, rc):
        self.e._check(rc)

    def begin(self, gt, prerank, rsamp):
        self._keep = (gt, prerank, rsamp)
        self._chk(self.e._lib.b3m_engine_blk_begin(self.e._h, self.prerate, self._largelcpthres, gt.data_ptr(), prerank.data_ptr(), rsamp.data_ptr()))

    def build_range(self, a0, a1, nblocks, L):
        t = C.c_uint32(0)
        self._chk(self.e._lib.b3m_engine_blk_build_range(self.e._h, a0, a1, nblocks, L.data_ptr(), C.byref(t)))
        return int(t.value)

    def chains(self, nr):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._chk(self.e._lib.b3m_engine_blk_chains(self.e._h, nr, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def zranks(self, a0, a1, r1, chl, nch, r0):
        self._chk(self.e._lib.b3m_engine_blk_zranks(self.e._h, a0, a1, r1, chl, nch, r0.data_ptr()))

    def gap(self, LA, a0, na, termA, r1, chl, nch, c_lo, c_hi, r0, gtnew, G):
        self._chk(self.e._lib.b3m_engine_blk_gap(self.e._h, LA.data_ptr(), a0, na, termA, r1, chl, nch, c_lo, c_hi, r0.data_ptr(),
                                                 gtnew.data_ptr(), G.data_ptr()))

    def merge(self, LA, na, termA, LR, nr, termR, a1, G, LM):
        t = C.c_uint32(0)
        self._chk(self.e._lib.b3m_engine_blk_merge(self.e._h, LA.data_ptr(), na, termA, LR.data_ptr(), nr, termR, a1, G.data_ptr(),
                                                   LM.data_ptr(), C.byref(t)))
        return int(t.value)

    def merge_samples(self, a0, a1, r1, G):
        self._chk(self.e._lib.b3m_engine_blk_merge_samples(self.e._h, a0, a1, r1, G.data_ptr()))

    def finish(self, L, term, q_lo, q_hi, sarate, isarate, bwtonly, numblocks):
        self._chk(self.e._lib.b3m_engine_blk_finish(self.e._h, L.data_ptr(), term, q_lo, q_hi, sarate, isarate, 1 if bwtonly else 0, numblocks))
        if bwtonly:
            return None, None
        i = self.e.info()
        p = [C.c_void_p() for _ in range(4)]
        self._chk(self.e._lib.b3m_engine_device_results(self.e._h, *[C.byref(x) for x in p]))
        sa = torch.as_tensor(_DevArray(p[2].value, i["nsa"], "<i8"), device=self.device)
        isa = torch.as_tensor(_DevArray(p[3].value, i["nisa"], "<i8"), device=self.device)
        return sa, isa


class ShardBuffers:
    """Zeroed device buffers one rank's suffix-range build writes at GLOBAL places (b3m_engine_shard_build)."""

    def __init__(self, engine, preisarate, sasamplingrate, isasamplingrate, bwtonly, device=None):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        n = engine.info()["n"]
        self.prerate = preisarate or engine.default_preisarate(bwtonly)
        self.n = n
        z = lambda k, dt: torch.zeros(k, dtype=dt, device=self.device)
        self.bwt = z(n + 16, torch.uint8)
        self.prerank = z((n + self.prerate - 1) // self.prerate, torch.int32)
        self.sa = None if bwtonly else z((n + sasamplingrate - 1) // sasamplingrate, torch.int64)
        self.isa = None if bwtonly else z((n + isasamplingrate - 1) // isasamplingrate, torch.int64)
        self.special = z(4, torch.int32)
        self.packed = None  # transport buffer of the slice exchange

    def zero_(self):
        for t in (self.bwt, self.prerank, self.sa, self.isa, self.special):
            if t is not None:
                t.zero_()

    def ptrs(self):
        p = lambda t: t.data_ptr() if t is not None else 0
        return p(self.bwt), p(self.prerank), p(self.sa), p(self.isa), p(self.special)

    def tensors(self):
        return [t for t in (self.bwt, self.prerank, self.sa, self.isa, self.special) if t is not None]


def build_sharded(engine, preisarate=0, sasamplingrate=32, isasamplingrate=262144, bwtonly=False, buffers=None, rank=None, world=None):
    """Suffix-range sharding: rank r sorts key range r of the replicated
    text; the slices combine by one sum-reduce to rank 0, whose engine then holds the complete results.
    Returns (ok, buffers): ok is False when some rank met repeats this path does not sort -- the
    caller then runs the block merge tree (build_distributed(..., strategy="merge"))."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    buf = buffers or ShardBuffers(engine, preisarate, sasamplingrate, isasamplingrate, bwtonly)
    if buffers is not None:
        if world > 1 and hasattr(engine, "shard_rows"):
            # reused buffers: only the summed (sparse) ones must start from zero, the dense slices are overwritten
            for t in (buf.prerank, buf.isa, buf.special):
                if t is not None:
                    t.zero_()
        else:
            buf.zero_()
    unres = engine.shard_build(rank, world, *buf.ptrs(), preisarate=buf.prerate, sasamplingrate=sasamplingrate,
                               isasamplingrate=isasamplingrate, bwtonly=bwtonly)
    flag = torch.tensor([unres], dtype=torch.int64, device=buf.device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
    if int(flag.item()) != 0:
        return False, buf
    if world > 1:
        # dense results (BWT rows, rank-sampled SA) are contiguous per rank: they travel as slices straight
        # into rank 0's buffers; the sparse ones (anchors, ISA samples, special rows) are summed -- every entry
        # is written by exactly one rank, the others hold 0
        rows = engine.shard_rows(world) if hasattr(engine, "shard_rows") else None
        if rows is not None:
            sr = sasamplingrate
            ops = []
            # BWT rows of a DNA text travel 2 bit per row (a quarter of rank 0's NVLink ingress)
            packed = None
            if hasattr(engine, "pack_rows"):
                if buf.packed is None:
                    buf.packed = torch.empty((buf.n + 3) // 4 + world, dtype=torch.uint8, device=buf.device)
                mine_lo, mine_hi = rows[rank], rows[rank + 1]
                if rank == 0 or engine.pack_rows(buf.bwt[mine_lo:].data_ptr(), mine_hi - mine_lo, buf.packed[mine_lo // 4 + rank:].data_ptr()):
                    packed = buf.packed
            flag2 = torch.tensor([1 if packed is not None else 0], dtype=torch.int64, device=buf.device)
            dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
            if int(flag2.item()) == 0:
                packed = None
            for p in range(1, world):
                lo, hi = rows[p], rows[p + 1]
                pieces = [packed[lo // 4 + p:lo // 4 + p + (hi - lo + 3) // 4] if packed is not None else buf.bwt[lo:hi]]
                if buf.sa is not None:
                    pieces.append(buf.sa[(lo + sr - 1) // sr:(hi + sr - 1) // sr])
                for t in pieces:
                    if t.numel() == 0:
                        continue
                    if rank == 0:
                        ops.append(dist.P2POp(dist.irecv, t, p))
                    elif rank == p:
                        ops.append(dist.P2POp(dist.isend, t, 0))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            if packed is not None and rank == 0:
                for p in range(1, world):
                    lo, hi = rows[p], rows[p + 1]
                    if hi > lo:
                        engine.unpack_rows(packed[lo // 4 + p:].data_ptr(), hi - lo, buf.bwt[lo:].data_ptr())
            sparse = [t for t in (buf.prerank, buf.isa, buf.special) if t is not None]
        else:
            sparse = buf.tensors()
        for t in sparse:
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        engine.shard_finish(world, *buf.ptrs())
    return True, buf


class DirectResults:
    """Result buffers of a sharded build that every rank writes directly: rank 0 owns them (b3m_dev_alloc),
    the other ranks map them through CUDA IPC, and the kernels of b3m_engine_shard_build store their BWT rows,
    anchors and samples into rank 0's HBM over NVLink.  No gather, no reduction: every entry is written by
    exactly one rank.  `mem` is a DeviceMemory-like object (alloc / export / open); the CPU tests plug in a
    shared-memory model of it."""

    def __init__(self, engine, preisarate, sasamplingrate, isasamplingrate, bwtonly, rank, world, mem=None):
        from .engine import DeviceMemory
        n = engine.info()["n"]
        self.n, self.rank, self.world = n, rank, world
        self.key = (n, sasamplingrate, isasamplingrate, bool(bwtonly))
        self.prerate = preisarate or engine.default_preisarate(bwtonly)
        self.mem = mem or DeviceMemory(engine.device)
        sizes = [("bwt", n + 16), ("prerank", 4 * ((n + self.prerate - 1) // self.prerate))]
        if not bwtonly:
            sizes += [("sa", 8 * ((n + sasamplingrate - 1) // sasamplingrate)), ("isa", 8 * ((n + isasamplingrate - 1) // isasamplingrate))]
        sizes.append(("special", 16))
        self.sizes = dict(sizes)
        self.p = {}
        handles = None
        if rank == 0:
            self.p = {k: self.mem.alloc(sz) for k, sz in sizes}
            handles = {k: self.mem.export(v) for k, v in self.p.items()}
        if world > 1:
            box = [handles]
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                self.p = {k: self.mem.open(h) for k, h in box[0].items()}

    def ptrs(self):
        g = lambda k: self.p.get(k, 0)
        return g("bwt"), g("prerank"), g("sa"), g("isa"), g("special")

    def extra(self, name, nbytes):
        """One more buffer owned by rank 0 and mapped by all (collective on first use)."""
        if name in self.p:
            return self.p[name]
        box = [None]
        if self.rank == 0:
            self.p[name] = self.mem.alloc(nbytes)
            box[0] = self.mem.export(self.p[name])
        if self.world > 1:
            dist.broadcast_object_list(box, src=0)
            if self.rank != 0:
                self.p[name] = self.mem.open(box[0])
        self.sizes[name] = nbytes
        return self.p[name]

    def close(self):
        for v in self.p.values():
            (self.mem.free if self.rank == 0 else self.mem.close)(v)
        self.p = {}
        if getattr(self, "sa_local", None):
            self.mem.free(self.sa_local)  # this rank's own copy of the sampled SA (build_xsharded, streamed samples)
            self.sa_local = None


def build_sharded_direct(engine, results, sasamplingrate=32, isasamplingrate=262144, bwtonly=False, device=None):
    """Suffix-range sharding with direct stores (DirectResults): rank r sorts key range r and writes its outputs
    into rank 0's buffers.  One all-reduce of the "suffixes left unresolved" counts on the build's stream is the
    only collective: it is the vote on the fast path and, being ordered behind every rank's kernels, the point
    after which rank 0 may read.  Returns False when the text needs the general path (block merge tree).
    The next build must not start before rank 0 has consumed the results (a barrier, or the next collective)."""
    rank, world = results.rank, results.world
    unres = engine.shard_build(rank, world, *results.ptrs(), preisarate=results.prerate, sasamplingrate=sasamplingrate,
                               isasamplingrate=isasamplingrate, bwtonly=bwtonly)
    if world > 1:
        flag = torch.tensor([unres], dtype=torch.int64, device=device or torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        unres = int(flag.item())
    if unres != 0:
        return False
    if rank == 0:
        engine.shard_adopt(world, *results.ptrs())
    return True


class XRecs:
    """The record arrays of a position-sharded build (b3m_engine_xshard_*): every rank owns one, sized for its key
    range, and maps all the others through CUDA IPC, so that the scatter kernel of rank p stores a record of rank
    q's key range straight into q's HBM."""

    def __init__(self, engine, rank, world, mem=None):
        from .engine import DeviceMemory
        n = engine.info()["n"]
        self.n, self.rank, self.world = n, rank, world
        self.cap = n // world + n // (4 * world) + (1 << 22)
        self.mem = mem or DeviceMemory(engine.device)
        self.own = self.mem.alloc(8 * self.cap)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, self.mem.export(self.own))
        self.ptrs = [self.own if q == rank else self.mem.open(handles[q]) for q in range(world)]

    def close(self):
        for q, p in enumerate(self.ptrs):
            (self.mem.free if q == self.rank else self.mem.close)(p)
        self.ptrs = []


def build_xsharded(engine, results, xrecs, sasamplingrate=32, isasamplingrate=262144, bwtonly=False, device=None):
    """Position-sharded build: every rank counts and scatters 1/world of the text positions,
    the records cross NVLink as the stores of the scatter kernel, every rank then sorts its own key range and stores
    its outputs into rank 0's buffers (DirectResults).  Collectives: one all-gather of the bin counts (16 KiB per
    rank), one all-reduce as the fence between scatter and sort, one all-reduce as vote + fence at the end.
    Returns None when the path does not apply to the text (the caller uses build_sharded_direct), False when the
    text needs the general path (block merge tree), True when rank 0's engine holds the results."""
    from .engine import B3MError
    rank, world = results.rank, results.world
    device = device or torch.device("cuda", torch.cuda.current_device())
    marks = getattr(results, "timeline", None)  # bench.py: a list that receives (phase, CUDA event) pairs of this build

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    tot = torch.zeros(2048, dtype=torch.int64, device=device)
    nb = engine.xshard_count(rank, world, tot.data_ptr(), preisarate=results.prerate, sasamplingrate=sasamplingrate,
                             isasamplingrate=isasamplingrate, bwtonly=bwtonly)
    if nb == 0:
        return None
    mark("count")
    allt = torch.empty(world * 2048, dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather_into_tensor(allt, tot)
    else:
        allt.copy_(tot)
    h = allt.cpu().numpy().view(np.uint64).reshape(world, 2048)[:, :nb]
    mark("totals all-gather")
    try:
        engine.xshard_scatter(h, xrecs.ptrs, [xrecs.cap] * world)
    except B3MError:
        return None  # a key range too large for its array: decided from the same counts on every rank
    mark("scatter")
    fence = torch.zeros(1, dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(fence, op=dist.ReduceOp.SUM)  # every rank's records have landed before anyone sorts
    mark("fence")
    # end to end: this rank's SA samples leave for the host while its finish kernel runs (fetch_distributed skips them then)
    host_sa = getattr(results, "stream_sa_host", 0)
    streamed = bool(host_sa) and not bwtonly and sasamplingrate >= 32 and hasattr(engine, "xshard_stream_sa")
    if streamed:
        if getattr(results, "sa_local", None) is None:
            results.sa_local = results.mem.alloc(results.sizes["sa"])
        engine.xshard_stream_sa(results.sa_local, host_sa)
    results.sa_streamed = False
    unres = engine.xshard_finish(xrecs.own, *results.ptrs())
    mark("local+finish")
    missed = 1 if (streamed and not engine.xshard_sa_delivered()) else 0  # a key range of too few sub-buckets does not stream
    if world > 1:
        flag = torch.tensor([unres, missed], dtype=torch.int64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        unres, missed = (int(x) for x in flag.tolist())
    mark("vote")
    if unres != 0:
        return False
    results.sa_streamed = streamed and missed == 0
    if rank == 0:
        engine.shard_adopt(world, *results.ptrs())
    return True


def gpu_numa_node(device_index):
    """NUMA node the GPU hangs off (sysfs numa_node of its PCI function), or -1 when the box does not say."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            return int(f.read().strip())
    except Exception:
        return -1


def _set_mempolicy(node):
    """Memory policy of the calling thread: pages it touches first are taken from `node` (MPOL_PREFERRED), or the
    default policy again for node < 0.  Returns False where the kernel / container does not allow the call."""
    try:
        libc = C.CDLL(None, use_errno=True)
        SYS_set_mempolicy = 238  # x86_64
        if node < 0:
            return libc.syscall(SYS_set_mempolicy, 0, None, 0) == 0
        mask = (C.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(SYS_set_mempolicy, 1, mask, 16 * 64 + 1) == 0
    except Exception:
        return False


def rank_slice(count, rank, world):
    """[lo, hi) of `count` result elements that rank `rank` sends to the host (fetch_distributed)."""
    per = ((count + world - 1) // world + 63) // 64 * 64
    return min(rank * per, count), min((rank + 1) * per, count)


class SharedHost:
    """A host buffer that every process of the job maps and page-locks: rank 0 creates a file under /dev/shm, all
    ranks map it (torch.from_file, shared), touch THEIR slice of it first -- with the memory policy set to the NUMA
    node of their GPU, so that the pages a rank's copy engine will write lie behind its own PCIe root and not across
    the socket interconnect -- and register the mapping with CUDA.  `width` = bytes per result element (the slices
    are those of fetch_distributed)."""

    def __init__(self, nbytes, tag, rank, world, width=1, device=None, directory="/dev/shm"):
        self.nbytes = max(int(nbytes), 1)
        self.t = None
        box = [None]
        if rank == 0:
            # a sparse file on a tmpfs without room would fault (SIGBUS) at the first touch: check before creating it
            import shutil
            import tempfile
            for d in (directory, tempfile.gettempdir()):
                try:
                    if os.path.isdir(d) and shutil.disk_usage(d).free > self.nbytes + (64 << 20):
                        box[0] = os.path.join(d, "b3m_%d_%s" % (os.getpid(), tag))
                        with open(box[0], "wb") as f:
                            f.truncate(self.nbytes)
                        break
                except OSError:
                    box[0] = None
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        if box[0] is None:
            raise RuntimeError("no room for a shared host buffer of %d bytes under %s" % (self.nbytes, directory))  # on every rank alike
        self.path = box[0]
        self.t = torch.from_file(self.path, shared=True, size=self.nbytes, dtype=torch.uint8)
        lo, hi = rank_slice(self.nbytes // width, rank, world)
        self.node = gpu_numa_node(torch.cuda.current_device() if device is None else device)
        self.policy = self.node >= 0 and _set_mempolicy(self.node)
        if hi > lo:
            self.t[lo * width:hi * width].zero_()  # first touch: these pages now exist, on this rank's node if the policy took
        if self.policy:
            _set_mempolicy(-1)
        if world > 1:
            dist.barrier()
        rc = torch.cuda.cudart().cudaHostRegister(self.t.data_ptr(), self.nbytes, 0)
        if int(rc) != 0:
            raise RuntimeError("cudaHostRegister failed (%s)" % rc)
        if world > 1:
            dist.barrier()
        if rank == 0:
            os.unlink(self.path)  # the mappings keep the memory alive

    def ptr(self):
        return self.t.data_ptr()

    def close(self):
        if self.t is not None:
            torch.cuda.cudart().cudaHostUnregister(self.t.data_ptr())
            self.t = None


def load_distributed(engine, host, inputtype, state, device=None):
    """The input file reaches the GPUs over ALL their PCIe links: rank r copies bytes [r*chunk, (r+1)*chunk) of the
    page-locked file image `host` (a uint8 tensor every rank holds or maps) into its slice of a device buffer, one
    all-gather over NVLink completes the buffer on every rank, and every rank decodes the text (K1).  Runs on
    torch's current stream, which must be the engine's.  `state` (a dict) keeps the device buffer between calls."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = host.numel()
    chunk = ((n + world - 1) // world + 255) // 256 * 256
    buf = state.get("in")
    if buf is None or buf.numel() != chunk * world:
        buf = state["in"] = torch.empty(chunk * world, dtype=torch.uint8, device=device or torch.device("cuda", torch.cuda.current_device()))
    lo, hi = min(rank * chunk, n), min((rank + 1) * chunk, n)
    if hi > lo:
        buf[lo:hi].copy_(host[lo:hi], non_blocking=True)
    if world > 1:
        dist.all_gather_into_tensor(buf, buf[rank * chunk:(rank + 1) * chunk])
    engine.load_device(buf.data_ptr(), n, inputtype)
    return hi - lo


def fetch_distributed(engine, state, host_words_ptr, host_sa_ptr, host_preisa_ptr=0, host_isa_ptr=0, device=None):
    """Results of a sharded pacterm build (build_distributed, strategy "shard") to the host over ALL PCIe links:
    rank 0 packs BWA's BWT words (K9) into a buffer the other ranks map, every rank pulls its slice of the words
    and of the sampled SA out of rank 0's HBM over NVLink (cudaMemcpyAsync on peer-mapped memory) and copies it to
    its place in the page-locked host buffers (SharedHost: the same memory in every process).  Anchors and ISA
    samples (n / 8192 and n / 262144 entries) leave through rank 0.  Two one-element all-reduces on the stream:
    "the words are packed" and "every slice has left rank 0's buffers" (the next build may overwrite them).
    Returns the bytes this rank sent to the host."""
    res = state.get("direct")
    if res is None:
        raise RuntimeError("fetch_distributed follows a sharded build_distributed with the same driver state")
    rank, world = res.rank, res.world
    dev = device or torch.device("cuda", torch.cuda.current_device())  # (a CPU device only in the gloo tests of the schedule)
    stream_ptr = engine.stream_ptr
    n = res.n
    nwords = (n - 1 + 15) >> 4
    nsa = res.sizes.get("sa", 0) // 8
    res.extra("bwa", 4 * nwords)
    marks = state.get("fetch_timeline")  # bench.py: a list that receives (phase, CUDA event) pairs

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    fence = state.get("fence")
    if fence is None:
        fence = state["fence"] = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        engine.pack_bwa(res.p["bwa"], 0, nwords)
    mark("pack words (rank 0)")
    if world > 1:
        dist.all_reduce(fence, op=dist.ReduceOp.SUM)
    mark("fence")
    sent = 0
    todo = []
    sa_there = bool(getattr(res, "sa_streamed", False)) and host_sa_ptr and host_sa_ptr == getattr(res, "stream_sa_host", 0)
    if sa_there and rank == 0 and nsa:
        res.mem.copy(host_sa_ptr, res.p["sa"], 8, stream_ptr)  # sample 0 of a terminated text is written by rank 0's engine, not by a finish kernel
    for key, count, width, hptr in (("bwa", nwords, 4, host_words_ptr), ("sa", nsa, 8, host_sa_ptr)):
        if not count or not hptr or (key == "sa" and sa_there):
            continue
        lo, hi = rank_slice(count, rank, world)
        if hi <= lo:
            continue
        per = rank_slice(count, 0, world)[1]
        nb = (hi - lo) * width
        src = res.p[key] + lo * width
        if rank != 0:
            stage = state.get("stage_" + key)
            if stage is None or stage.numel() < nb:
                stage = state["stage_" + key] = torch.empty(per * width, dtype=torch.uint8, device=dev)
            res.mem.copy(stage.data_ptr(), src, nb, stream_ptr)  # NVLink: rank 0's HBM -> mine
            src = stage.data_ptr()
        todo.append((hptr + lo * width, src, nb))
    mark("pull slices over NVLink")
    for dst, src, nb in todo:
        res.mem.copy(dst, src, nb, stream_ptr)                     # my PCIe link
        sent += nb
    mark("slices to the host")
    if rank == 0 and (host_preisa_ptr or host_isa_ptr):
        engine.fetch_ptrs(0, host_preisa_ptr, 0, host_isa_ptr)
    mark("anchors + ISA (rank 0)")
    if world > 1:
        dist.all_reduce(fence, op=dist.ReduceOp.SUM)
    mark("fence 2")
    return sent


def build_distributed(engine, local_blocks=1, preisarate=0, sasamplingrate=32, isasamplingrate=262144, bwtonly=False,
                      largelcpthres=16384, driver=None, strategy="auto"):
    """Every rank has loaded the same text into `engine`; after the call rank 0's engine holds the
    complete results (fetch / write_bwt as after a single-GPU build).
    strategy: "shard" = suffix-range sharding only (raises if the text needs the general path),
    "merge" = the reference's block merge tree over NCCL, "auto" = shard, then merge if needed.
    `driver` caches the process groups / buffers between calls: pass back the first return value."""
    state = driver if isinstance(driver, dict) else {"merge": driver, "shard": None, "direct": None}
    if dist.get_backend() == "nccl" and not engine.stream_ptr:
        # the engine's kernels and torch's collectives must share one stream: nothing else orders them
        raise ValueError("build_distributed needs an Engine created on a torch stream: Engine(device, torch.cuda.Stream().cuda_stream)")
    if strategy in ("auto", "shard"):
        with torch.cuda.stream(torch.cuda.ExternalStream(engine.stream_ptr)) if engine.stream_ptr else _null():
            if hasattr(engine, "shard_adopt"):
                # every rank stores straight into rank 0's buffers (CUDA IPC)
                res = state.get("direct")
                key = (engine.info()["n"], sasamplingrate, isasamplingrate, bool(bwtonly))
                if res is not None and (res.key != key or (preisarate and res.prerate != preisarate)):
                    res.close()
                    res = None
                if res is None:
                    res = DirectResults(engine, preisarate, sasamplingrate, isasamplingrate, bwtonly, dist.get_rank(), dist.get_world_size())
                state["direct"] = res
                ok = None
                if hasattr(engine, "xshard_count") and dist.get_world_size() <= 16:
                    xr = state.get("xrecs")
                    if xr is not None and xr.n != key[0]:
                        xr.close()
                        xr = None
                    if xr is None:
                        xr = XRecs(engine, dist.get_rank(), dist.get_world_size())
                    state["xrecs"] = xr
                    res.stream_sa_host = state.get("stream_sa_host", 0)  # set by the caller between builds (a SharedHost pointer)
                    ok = build_xsharded(engine, res, xr, sasamplingrate, isasamplingrate, bwtonly)
                if ok is None:
                    ok = build_sharded_direct(engine, res, sasamplingrate, isasamplingrate, bwtonly)
            else:
                ok, state["shard"] = build_sharded(engine, preisarate, sasamplingrate, isasamplingrate, bwtonly, buffers=state["shard"])
            torch.cuda.current_stream().synchronize()
        if ok:
            return state, {"strategy": "shard"}
        if strategy == "shard":
            raise RuntimeError("suffix-range sharding left suffixes unresolved; use strategy='merge'")
    ops = EngineOps(engine, preisarate, largelcpthres)
    driver = state["merge"]
    drv = driver or DistBuild(ops)
    drv.ops = ops
    with torch.cuda.stream(torch.cuda.ExternalStream(engine.stream_ptr)) if engine.stream_ptr else _null():
        res = drv.build(local_blocks, sasamplingrate, isasamplingrate, bwtonly)
        torch.cuda.current_stream().synchronize()
    state["merge"] = drv
    res["strategy"] = "merge"
    return state, res


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
