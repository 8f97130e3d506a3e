"""GPU parity, multi-block: block sort with look-ahead (A4/A5) + gap arrays (A7) + gap-driven
merge (A8) through the C ABI must give the same BWT, anchors, sampled SA and ISA as the CPU
oracle -- and therefore the same as the single-block build (build plan step 5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from bwtb3m_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def check(eng, oracle, data, itype, text, numblocks, rates=(16, 4, 8), gapmode="auto"):
    sa = oracle.sa_circular(text)
    bwt, isa = oracle.bwt_from_sa(text, sa)
    pre, sar, isar = rates
    eng.load_host(data, itype)
    eng.build(numblocks=numblocks, preisarate=pre, sasamplingrate=sar, isasamplingrate=isar, gapmode=gapmode)
    res, info = eng.fetch(), eng.info()
    assert info["n"] == text.size
    assert np.array_equal(res["bwt"], bwt), "BWT differs (numblocks=%d)" % numblocks
    assert np.array_equal(res["preisa"][:, 0], isa[::pre].astype(np.uint64))
    assert np.array_equal(res["sa"], sa[::sar].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::isar].astype(np.uint64))
    return info


@pytest.mark.parametrize("numblocks", [2, 3, 4, 7, 8, 16])
@pytest.mark.parametrize("seed,n,sigma", [(1, 1000, 4), (2, 70001, 4), (3, 30000, 256), (4, 5000, 2), (5, 20000, 20), (6, 17, 3)])
def test_blocks_bytestream(eng, oracle, seed, n, sigma, numblocks):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, sigma, size=n, dtype=np.uint8)
    info = check(eng, oracle, t, "bytestream", t, numblocks)
    assert info["numblocks"] >= 2 and info["gap_lf_steps"] > 0


@pytest.mark.parametrize("numblocks", [2, 3, 5, 8])
@pytest.mark.parametrize("seed,l", [(11, 8), (12, 1001), (13, 65536), (14, 250003), (15, 3), (16, 16), (17, 15)])
@pytest.mark.parametrize("itype", ["pac", "pacterm"])
def test_blocks_pac(eng, oracle, seed, l, itype, numblocks):
    rng = np.random.default_rng(seed)
    bases = rng.integers(0, 4, size=l, dtype=np.uint8)
    if itype == "pac" and l == 16:
        bases[0] = (bases[1] + 1) % 4  # keep the text primitive
    pac = oracle.encode_pac(bases)
    t = oracle.decode_pac(pac.tobytes(), term=(itype == "pacterm"))
    check(eng, oracle, pac, itype, t, numblocks)


@pytest.mark.parametrize("gapmode", ["list", "atomic"])
@pytest.mark.parametrize("numblocks", [2, 3, 5, 16])
@pytest.mark.parametrize("case", ["pacterm-300007", "pac-65537", "bytes256-40001", "bytes3-9000", "pacterm-777", "pacterm-5"])
def test_blocks_gap_counting_modes(eng, oracle, case, numblocks, gapmode):
    """K5 counts the gap array either by atomic adds from the chains or from a list of the chains' ranks partitioned by
    one radix pass (what gap arrays beyond the L2 size take): forced on small texts, both equal the oracle."""
    kind, n = case.split("-")
    n = int(n)
    rng = np.random.default_rng(n + numblocks)
    if kind.startswith("bytes"):
        t = rng.integers(0, int(kind[5:]), size=n, dtype=np.uint8)
        data, itype = t, "bytestream"
    else:
        data, itype = oracle.encode_pac(rng.integers(0, 4, size=n, dtype=np.uint8)), kind
        t = oracle.decode_pac(data.tobytes(), term=(kind == "pacterm"))
    info = check(eng, oracle, data, itype, t, numblocks, gapmode=gapmode)
    assert info["gap_lf_steps"] > 0


def test_blocks_terminator_alone_in_last_block(eng, oracle):
    """n = l+1 with l divisible by the block count: the last block holds only the terminator."""
    rng = np.random.default_rng(31)
    bases = rng.integers(0, 4, size=4096, dtype=np.uint8)
    pac = oracle.encode_pac(bases)
    t = oracle.decode_pac(pac.tobytes(), term=True)
    for nb in (2, 4, 4097 // 2):
        check(eng, oracle, pac, "pacterm", t, nb)


@pytest.mark.parametrize("itype", ["bytestream", "pacterm"])
def test_blocks_repetitive_large_lcp(eng, oracle, itype):
    """Mutated copies: lcpnext runs into the thousands, above a small largelcpthres, so the
    bounded-then-exact escape path is taken (A4)."""
    rng = np.random.default_rng(41)
    base = rng.integers(0, 4, size=6000, dtype=np.uint8)
    parts = []
    for c in range(8):
        x = base.copy()
        pos = rng.integers(0, x.size, size=2)
        x[pos] = (x[pos] + 1) % 4
        parts.append(x)
    bases = np.concatenate(parts)
    if itype == "pacterm":
        data = oracle.encode_pac(bases)
        t = oracle.decode_pac(data.tobytes(), term=True)
    else:
        data = t = bases
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    for nb in (2, 5, 8):
        eng.load_host(data, itype)
        eng.build(numblocks=nb, preisarate=64, sasamplingrate=32, isasamplingrate=64, largelcpthres=256)
        res, info = eng.fetch(), eng.info()
        assert info["max_lcpnext"] >= 256
        assert np.array_equal(res["bwt"], bwt)
        assert np.array_equal(res["sa"], sa[::32].astype(np.uint64))
        assert np.array_equal(res["isa"], isa[::64].astype(np.uint64))


def test_blocks_skewed_gaps(eng, oracle):
    """A^k followed by random text: one gap of the top merge holds almost all of R (big-gap path)."""
    rng = np.random.default_rng(51)
    t = np.concatenate([rng.integers(1, 4, size=20000, dtype=np.uint8), np.zeros(20000, dtype=np.uint8),
                        rng.integers(0, 4, size=100, dtype=np.uint8)])
    check(eng, oracle, t, "bytestream", t, 2)
    check(eng, oracle, t, "bytestream", t, 4)


def test_blocks_equal_single_block_2mbp(eng, oracle):
    """Size-independent property at a larger size: multi-block == single-block, and the restated
    checkbwt verifier accepts the result."""
    from bwtb3m_b200 import workloads
    pac = workloads.random_pac(2_000_003, 7)
    eng.load_host(pac, "pacterm")
    eng.build(numblocks=1, sasamplingrate=32, isasamplingrate=1024)
    one = eng.fetch()
    for nb in (2, 8):
        eng.load_host(pac, "pacterm")
        eng.build(numblocks=nb, sasamplingrate=32, isasamplingrate=1024)
        res, info = eng.fetch(), eng.info()
        assert info["numblocks"] == nb
        for k in ("bwt", "preisa", "sa", "isa"):
            assert np.array_equal(res[k], one[k]), k
    t = oracle.decode_pac(pac.tobytes(), term=True)
    rc, checked = oracle.checkbwt(t, one["bwt"], one["preisa"], numthreads=8)
    assert rc == 1 and checked == t.size
