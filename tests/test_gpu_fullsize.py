"""GPU, BASELINE.json's full sizes: size-independent properties of the results where the oracle
would take too long.
  * the reference's own method for the sampled SA/ISA (LF walk over the finished BWT from the
    anchors, /root/reference/src/hwtPreIsaToIsa.cpp:114-161) must reproduce, value for value, the
    samples the sort emitted directly: one wrong BWT symbol or anchor derails a chain;
  * a 2-block build (gap array + merge, the reference's algorithm) equals the 1-block build;
  * symbol counts of the BWT equal those of the text; the rank-sampled SA is strictly increasing
    in suffix order (checked on the text for a few thousand neighbouring samples)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _decode(itype, data, lo, hi):
    """symbols [lo,hi) of the text in the reference's symbol space (pacterm: bases+1)"""
    if itype == "bytestream":
        return data[lo:hi].astype(np.int64)
    idx = np.arange(lo, hi, dtype=np.int64)
    return ((data[idx >> 2] >> ((~idx & 3) << 1)) & 3).astype(np.int64) + 1


def _sorted_spot_check(itype, data, n, sa, rate, picks=4000, depth=256):
    rng = np.random.default_rng(0)
    ks = rng.integers(0, sa.size - 1, size=picks)
    for k in ks:
        p, q = int(sa[k]), int(sa[k + 1])
        m = min(depth, n - 1 - max(p, q))  # stay in front of the terminator / text end
        if m <= 0:
            continue
        a, b = _decode(itype, data, p, p + m), _decode(itype, data, q, q + m)
        d = np.nonzero(a != b)[0]
        if d.size:
            assert a[d[0]] < b[d[0]], "suffix order violated at sampled ranks %d, %d" % (k * rate, (k + 1) * rate)
        elif itype == "pacterm" and max(p, q) + m >= n - 1:
            assert p > q, "equal up to the terminator: the shorter suffix must come first"
        # else: longer common prefix than `depth` symbols (repetitive input) -- not decided here


def _run(workload, scale, with_blocks=True, picks=4000, depth=256):
    import torch
    from bwtb3m_b200 import Engine, workloads
    free, _ = torch.cuda.mem_get_info()
    itype, data, nsym = workloads.make(workload, scale)
    if free < 36 * nsym + (2 << 30):
        pytest.skip("not enough free device memory for %s" % workload)
    eng = Engine(0)
    try:
        eng.load_host(data, itype)
        eng.build(sampling="auto")
        i1 = eng.info()
        a = eng.fetch()
        eng.load_host(data, itype)
        eng.build(sampling="walk")
        i2 = eng.info()
        b = eng.fetch(bwt=False)
        assert i1["walk_lf_steps"] == 0 and i2["walk_lf_steps"] == i1["n"]
        assert np.array_equal(a["sa"], b["sa"]) and np.array_equal(a["isa"], b["isa"]) and np.array_equal(a["preisa"], b["preisa"])
        if with_blocks:
            eng.load_host(data, itype)
            eng.build(numblocks=2, preisarate=i1["preisarate"])
            c = eng.fetch()
            assert np.array_equal(a["bwt"], c["bwt"]) and np.array_equal(a["sa"], c["sa"]) and np.array_equal(a["isa"], c["isa"])
            assert np.array_equal(a["preisa"], c["preisa"])
            del c
    finally:
        eng.close()
    n = i1["n"]
    counts = np.bincount(a["bwt"], minlength=256)
    assert {s: int(v) for s, v in enumerate(counts) if v} == i1["hist"]
    assert a["sa"].size == (n + 31) // 32 and int(a["sa"].max()) < n and np.unique(a["sa"]).size == a["sa"].size
    _sorted_spot_check(itype, data, n, a["sa"], 32, picks, depth)
    return i1


def test_cfg3_full_size():
    """3.1 Gbp pacterm genome (BASELINE config 3, the bench workload)."""
    info = _run("cfg3", 1.0)
    assert info["n"] == 3_100_000_001 and info["sort_unresolved0"] == 0


def test_cfg5_quarter_size_bytes():
    """Byte alphabet, circular, no terminator (BASELINE config 5 at a quarter of its size)."""
    info = _run("cfg5", 0.25)
    assert info["sigma"] == 256


def test_cfg4_repetitive_tenth_size():
    """64 mutated copies (BASELINE config 4 at a tenth of its size): LCPs in the thousands, the
    prefix-doubling rounds and the large-LCP block path must agree with the direct sampling."""
    info = _run("cfg4", 0.1, picks=400, depth=40000)
    assert info["sort_unresolved0"] > 0 and info["sort_rounds"] > 4


def _oracle_checkbwt_full(workload, oracle):
    """The reference's own verifier (checkbwt, /root/reference/src/checkbwt.cpp:126-243, restated in
    oracle/b3m_oracle.c:orc_checkbwt) on the GPU's BWT + (rank,pos) anchors at the FULL size of the config:
    an LF walk over the whole BWT on the host cores that compares every symbol with the circularly reversed
    text and must arrive at every anchor's rank -- a complete proof of the symbol stream."""
    import os
    import torch
    from bwtb3m_b200 import Engine, workloads
    free, _ = torch.cuda.mem_get_info()
    itype, data, nsym = workloads.make(workload, 1.0)
    if free < 36 * nsym + (2 << 30):
        pytest.skip("not enough free device memory for %s" % workload)
    eng = Engine(0)
    try:
        eng.load_host(data, itype)
        eng.build()
        info = eng.info()
        res = eng.fetch()
    finally:
        eng.close()
    text = oracle.decode_pac(data.tobytes(), term=True) if itype == "pacterm" else np.ascontiguousarray(data)
    n = info["n"]
    assert text.size == n and res["bwt"].size == n
    rc, checked = oracle.checkbwt(text, res["bwt"], res["preisa"], numthreads=os.cpu_count() or 8)
    assert rc == 1, "the reference's verifier rejects the GPU BWT of %s" % workload
    assert checked == n
    return info


def test_cfg3_full_size_oracle_checkbwt(oracle):
    info = _oracle_checkbwt_full("cfg3", oracle)
    assert info["n"] == 3_100_000_001


def test_cfg5_full_size_oracle_checkbwt(oracle):
    info = _oracle_checkbwt_full("cfg5", oracle)
    assert info["n"] == 1_000_000_000 and info["sigma"] == 256


def test_cfg4_full_size_oracle_checkbwt(oracle):
    info = _oracle_checkbwt_full("cfg4", oracle)
    assert info["n"] == 3_200_000_001 and info["sort_unresolved0"] > 0
