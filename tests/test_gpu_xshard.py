"""GPU: the position-sharded build (b3m_engine_xshard_count / _scatter / _finish, the multi-GPU path of the MSD
sorter) on ONE device: one engine per part, the record arrays and the result buffers are plain device buffers that
every part writes, the exchange of the counts is a host concatenation.  The result must equal the oracle exactly as
a single-block build does."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def xsharded(data, itype, nparts, **kw):
    import torch
    from bwtb3m_b200 import Engine, multigpu
    engs = [Engine(0) for _ in range(nparts)]
    try:
        for e in engs:
            e.load_host(data, itype)
        res = multigpu.DirectResults(engs[0], kw.get("preisarate", 0), kw["sasamplingrate"], kw["isasamplingrate"], False, 0, 1)
        n = engs[0].info()["n"]
        cap = n // nparts + n // (4 * nparts) + (1 << 16)
        recs = [torch.full((cap,), -1, dtype=torch.int64, device="cuda") for _ in range(nparts)]
        tots = [torch.zeros(2048, dtype=torch.int64, device="cuda") for _ in range(nparts)]
        nb = 0
        for p, e in enumerate(engs):
            nb = e.xshard_count(p, nparts, tots[p].data_ptr(), preisarate=res.prerate, sasamplingrate=kw["sasamplingrate"],
                                isasamplingrate=kw["isasamplingrate"], sortpath=kw.get("sortpath", "auto"))
        if nb == 0:
            return None, None
        torch.cuda.synchronize()
        allt = torch.stack(tots).cpu().numpy().view(np.uint64)[:, :nb]
        for e in engs:
            e.xshard_scatter(allt, [r.data_ptr() for r in recs], [cap] * nparts)
        torch.cuda.synchronize()
        unres = 0
        for p, e in enumerate(engs):
            unres += e.xshard_finish(recs[p].data_ptr(), *res.ptrs())
        torch.cuda.synchronize()
        if unres:
            return None, unres
        engs[0].shard_adopt(nparts, *res.ptrs())
        out = engs[0].fetch()
        bwa = engs[0].fetch_bwa() if itype == "pacterm" else None
        res.close()
        return out, bwa
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("nparts", [1, 2, 3, 8])
@pytest.mark.parametrize("itype,n", [("pacterm", 200_003), ("pac", 150_000), ("pacterm", 3_000_001), ("bytestream", 100_000)])
def test_xshard_equals_oracle(oracle, itype, n, nparts):
    rng = np.random.default_rng(n + nparts)
    bases = rng.integers(0, 4, size=n, dtype=np.uint8)
    if itype == "bytestream":
        data, t = bases, bases
    else:
        data = oracle.encode_pac(bases)
        t = oracle.decode_pac(data.tobytes(), term=(itype == "pacterm"))
    res, bwa = xsharded(data, itype, nparts, preisarate=64, sasamplingrate=8, isasamplingrate=32, sortpath="msd")
    assert res is not None
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["preisa"][:, 0], isa[::64].astype(np.uint64))
    assert np.array_equal(res["sa"], sa[::8].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::32].astype(np.uint64))
    if itype == "pacterm":
        assert bwa[1] == int(isa[0]) and bwa[3] == t.size - 1


def test_xshard_not_for_byte_alphabets():
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, size=100_000, dtype=np.uint8)
    res, _ = xsharded(data, "bytestream", 2, sasamplingrate=8, isasamplingrate=32)
    assert res is None


def test_xshard_reports_repeats(oracle):
    rng = np.random.default_rng(6)
    u = rng.integers(0, 4, size=50_000, dtype=np.uint8)
    bases = np.concatenate([u, rng.integers(0, 4, size=1000, dtype=np.uint8), u])
    data = oracle.encode_pac(bases)
    res, unres = xsharded(data, "pacterm", 2, sasamplingrate=8, isasamplingrate=32, sortpath="msd")
    assert res is None and unres > 0
