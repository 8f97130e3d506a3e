"""GPU parity of the MSD suffix sorter (csrc/msd.cuh; sortpath="msd", the default for texts of 2^16 symbols
or more over at most four codes) against the CPU oracle, and against the LSD sorter on the same inputs:
bit-exact BWT, anchors, sampled SA and ISA.  The forced mode runs the path on texts far smaller than the
automatic threshold, so that every branch (partial tiles, tiles at both ends of the text, sub-buckets too
large for one CTA, local digits too crowded to be compared, ties beyond the bits a record carries, suffixes
that reach the terminator) is reached at sizes the oracle sorts in seconds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from bwtb3m_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def run(eng, data, inputtype, **kw):
    eng.load_host(data, inputtype)
    eng.build(**kw)
    return eng.fetch(), eng.info()


def check_all(oracle, res, t, prerate, sarate, isarate):
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["preisa"][:, 0], isa[::prerate].astype(np.uint64))
    assert np.array_equal(res["sa"], sa[::sarate].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::isarate].astype(np.uint64))


def make(oracle, itype, bases):
    if itype == "bytestream":
        return bases, bases
    data = oracle.encode_pac(bases)
    return data, oracle.decode_pac(data.tobytes(), term=(itype == "pacterm"))


@pytest.mark.parametrize("l", [64, 65, 100, 1001, 8191, 8192, 8193, 16385, 65536, 250_003, 1_000_003])
@pytest.mark.parametrize("itype", ["pacterm", "pac", "bytestream"])
def test_random_forced_msd(eng, oracle, itype, l):
    rng = np.random.default_rng(l)
    data, t = make(oracle, itype, rng.integers(0, 4, size=l, dtype=np.uint8))
    res, info = run(eng, data, itype, preisarate=16, sasamplingrate=4, isasamplingrate=8, sortpath="msd")
    assert info["radix_passes"] == 2  # the two MSD levels ran, not the four LSD passes
    check_all(oracle, res, t, 16, 4, 8)


@pytest.mark.parametrize("itype,l", [("pacterm", 5_000_011), ("pac", 3_000_000), ("bytestream", 2_000_003)])
def test_random_auto_takes_msd(eng, oracle, itype, l):
    rng = np.random.default_rng(l)
    data, t = make(oracle, itype, rng.integers(0, 4, size=l, dtype=np.uint8))
    res, info = run(eng, data, itype, preisarate=64, sasamplingrate=32, isasamplingrate=128)
    assert info["radix_passes"] == 2 and info["sort_unresolved0"] == 0
    check_all(oracle, res, t, 64, 32, 128)
    res2, info2 = run(eng, data, itype, preisarate=64, sasamplingrate=32, isasamplingrate=128, sortpath="lsd")
    assert info2["radix_passes"] >= 3
    for k in ("bwt", "preisa", "sa", "isa"):
        assert np.array_equal(res[k], res2[k])


@pytest.mark.parametrize("sigma,l", [(2, 300_000), (3, 200_001), (1, 5000)])
@pytest.mark.parametrize("itype", ["pacterm", "bytestream"])
def test_small_alphabets(eng, oracle, itype, sigma, l):
    """Few distinct symbols: crowded sub-buckets and local digits, ties far beyond the record's bits (sigma 1: one
    sub-bucket holds everything; terminated text only, a circular text of one symbol has no defined SA)."""
    if sigma == 1 and itype != "pacterm":
        pytest.skip("periodic circular text: order of equal rotations is undefined")
    rng = np.random.default_rng(sigma * 7 + l)
    data, t = make(oracle, itype, rng.integers(0, sigma, size=l, dtype=np.uint8))
    res, info = run(eng, data, itype, preisarate=16, sasamplingrate=4, isasamplingrate=8, sortpath="msd")
    check_all(oracle, res, t, 16, 4, 8)


@pytest.mark.parametrize("itype", ["pacterm", "bytestream"])
def test_long_runs(eng, oracle, itype):
    """Runs of one symbol of up to 30000: sub-buckets above the capacity of a finish CTA (kept as one
    unresolved group), crowded local digits, then prefix doubling from a short common prefix."""
    rng = np.random.default_rng(78)
    parts = []
    for k in range(30):
        parts.append(np.full(int(rng.integers(50, 30000)), k % 4, dtype=np.uint8))
        parts.append(rng.integers(0, 4, size=int(rng.integers(1, 2000)), dtype=np.uint8))
    data, t = make(oracle, itype, np.concatenate(parts))
    res, info = run(eng, data, itype, preisarate=16, sasamplingrate=4, isasamplingrate=8, sortpath="msd")
    assert info["sort_unresolved0"] > 0 and info["sort_rounds"] > 1
    check_all(oracle, res, t, 16, 4, 8)


def test_mutated_copies(eng, oracle):
    """cfg4 in small: 16 copies with substitutions at rate 1e-3; every suffix ties beyond the second key."""
    from bwtb3m_b200 import workloads
    data = workloads.repetitive_pac(16, 40_000, 9)
    t = oracle.decode_pac(data.tobytes(), term=True)
    res, info = run(eng, data, "pacterm", preisarate=64, sasamplingrate=32, isasamplingrate=64, sortpath="msd")
    assert info["sort_unresolved0"] > t.size // 2
    check_all(oracle, res, t, 64, 32, 64)


@pytest.mark.parametrize("copies,base_len,rate", [(300, 3_000, 1e-3), (257, 2_500, 3e-3), (40, 20_000, 1e-4), (600, 1_200, 2e-2)])
def test_doubling_rounds_tile_and_radix(eng, oracle, copies, base_len, rate):
    """Prefix-doubling rounds (sufsort.cu k_dbl_tile): groups of up to 256 tied suffixes are split inside a CTA,
    larger ones go through the radix round; texts whose groups lie on both sides of that limit and cross it as the
    rounds split them.  The order must equal the oracle's and the one of the radix-only rounds (sortpath="lsd")."""
    from bwtb3m_b200 import workloads
    data = workloads.repetitive_pac(copies, base_len, copies + 1, rate=rate)
    t = oracle.decode_pac(data.tobytes(), term=True)
    res, info = run(eng, data, "pacterm", preisarate=64, sasamplingrate=32, isasamplingrate=64, sortpath="msd")
    assert info["sort_unresolved0"] > t.size // 2 and info["sort_rounds"] > 2
    check_all(oracle, res, t, 64, 32, 64)
    res2, info2 = run(eng, data, "pacterm", preisarate=64, sasamplingrate=32, isasamplingrate=64, sortpath="lsd")
    for k in ("bwt", "preisa", "sa", "isa"):
        assert np.array_equal(res[k], res2[k])


def test_doubling_rounds_circular_bytes(eng, oracle):
    """The same rounds on a circular text (bytestream ACGT codes): ranks ahead wrap around the end."""
    rng = np.random.default_rng(5)
    base = rng.integers(0, 4, size=9_000, dtype=np.uint8)
    parts = []
    for c in range(70):
        x = base.copy()
        pos = rng.integers(0, base.size, size=12)
        x[pos] = (x[pos] + 1) & 3
        parts.append(x)
    parts.append(rng.integers(0, 4, size=777, dtype=np.uint8))  # makes the circular text primitive
    data, t = make(oracle, "bytestream", np.concatenate(parts))
    res, info = run(eng, data, "bytestream", preisarate=64, sasamplingrate=32, isasamplingrate=64, sortpath="msd")
    assert info["sort_rounds"] > 2
    check_all(oracle, res, t, 64, 32, 64)


def test_tail_compares_by_length(eng, oracle):
    """pacterm: suffixes that reach the terminator inside the record's bits or the second key."""
    for l in (64, 65, 79, 80, 81, 95, 96, 97, 128, 200):
        for fill in (0, 3):
            bases = np.full(l, fill, dtype=np.uint8)
            bases[l // 2] = (fill + 1) % 4
            data = oracle.encode_pac(bases)
            t = oracle.decode_pac(data.tobytes(), term=True)
            res, info = run(eng, data, "pacterm", preisarate=1, sasamplingrate=1, isasamplingrate=1, sortpath="msd")
            check_all(oracle, res, t, 1, 1, 1)


@pytest.mark.parametrize("nparts", [1, 2, 3, 8])
@pytest.mark.parametrize("itype,n", [("pacterm", 200_003), ("pac", 150_000), ("pacterm", 2_000_000)])
def test_shards_msd(eng, oracle, itype, n, nparts):
    """Suffix-range sharding over the level-1 buckets of the MSD path, every key range on this one device."""
    import torch
    from bwtb3m_b200 import multigpu
    rng = np.random.default_rng(n + nparts)
    data, t = make(oracle, itype, rng.integers(0, 4, size=n, dtype=np.uint8))
    eng.load_host(data, itype)
    buf = multigpu.ShardBuffers(eng, 64, 8, 32, False)
    unres = 0
    for part in range(nparts):
        unres += eng.shard_build(part, nparts, *buf.ptrs(), preisarate=buf.prerate, sasamplingrate=8, isasamplingrate=32, sortpath="msd")
    torch.cuda.synchronize()
    assert unres == 0
    assert eng.info()["radix_passes"] == 2
    eng.shard_finish(nparts, *buf.ptrs())
    check_all(oracle, eng.fetch(), t, 64, 8, 32)


def test_streamed_results_msd(oracle):
    """host_sa / host_bwa delivery during the finish kernel's chunks."""
    import torch
    from bwtb3m_b200 import Engine
    rng = np.random.default_rng(43)
    e = Engine(0)
    try:
        bases = rng.integers(0, 4, size=1_500_003, dtype=np.uint8)
        pac = oracle.encode_pac(bases)
        e.load_host(pac, "pacterm")
        n = bases.size + 1
        host = torch.full(((n + 31) // 32,), -1, dtype=torch.int64).pin_memory()
        hbwa = torch.full(((n - 1 + 15) // 16,), -1, dtype=torch.int32).pin_memory()
        e.build(sasamplingrate=32, isasamplingrate=64, host_sa_ptr=host.data_ptr(), host_bwa_ptr=hbwa.data_ptr())
        info = e.info()
        assert info["radix_passes"] == 2 and info["sort_unresolved0"] == 0
        res = e.fetch()
        assert np.array_equal(host.numpy().astype(np.uint64), res["sa"])
        words, primary, l2, seq_len = e.fetch_bwa()
        assert np.array_equal(hbwa.numpy().view(np.uint32), words)
        t = oracle.decode_pac(pac.tobytes(), term=True)
        sa = oracle.sa_circular(t)
        bwt, isa = oracle.bwt_from_sa(t, sa)
        assert np.array_equal(res["sa"], sa[::32].astype(np.uint64)) and np.array_equal(res["bwt"], bwt)
        assert primary == int(isa[0])
    finally:
        e.close()
