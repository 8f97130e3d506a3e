"""GPU parity, single block: the CUDA path (through the C ABI, host buffers) against the CPU
oracle on the same seeded inputs; bit-exact BWT, anchors, sampled SA and ISA."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KATS = [
    (b"abbab#", b"bb#aba", [2, 5, 4, 1, 3, 0]),
    (b"banana", b"nnbaaa", [3, 2, 5, 1, 4, 0]),
    (b"mississippi", b"pssmipissii", [4, 3, 10, 8, 2, 9, 7, 1, 6, 5, 0]),
    (b"ACGTACGTTGCA", b"CATGAATCCGTG", [1, 4, 7, 9, 2, 5, 8, 11, 10, 6, 3, 0]),
]


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


@pytest.mark.parametrize("text,bwt,isa", KATS)
def test_kat_bytestream(eng, text, bwt, isa):
    res, info = run(eng, text, "bytestream", preisarate=1, sasamplingrate=1, isasamplingrate=1)
    assert res["bwt"].tobytes() == bwt
    assert res["preisa"][:, 0].tolist() == isa
    assert res["isa"].tolist() == isa
    sa = [0] * len(isa)
    for p, r in enumerate(isa):
        sa[r] = p
    assert res["sa"].tolist() == sa


def test_kat_pacterm(eng):
    res, info = run(eng, bytes([0x1B, 0xE4, 0x00, 0x00]), "pacterm", preisarate=1, sasamplingrate=1, isasamplingrate=1)
    assert info["n"] == 9
    assert res["bwt"].tolist() == [1, 2, 0, 3, 1, 4, 2, 4, 3]
    assert res["sa"].tolist() == [8, 7, 0, 6, 1, 5, 2, 4, 3]
    assert res["isa"][0] == 2
    assert info["hist"] == {0: 1, 1: 2, 2: 2, 3: 2, 4: 2}


@pytest.mark.parametrize("seed,n,sigma", [(1, 1000, 4), (2, 70001, 4), (3, 30000, 256), (4, 5000, 2), (5, 20000, 20),
                                          (6, 300000, 4), (7, 1, 4), (8, 2, 4), (9, 17, 3)])
def test_random_bytestream_vs_oracle(eng, oracle, seed, n, sigma):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, sigma, size=n, dtype=np.uint8)
    if n == 2:
        t[:] = [1, 0]
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    res, info = run(eng, t, "bytestream", preisarate=64, sasamplingrate=32, isasamplingrate=128)
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["preisa"][:, 0], isa[::64].astype(np.uint64))
    assert np.array_equal(res["preisa"][:, 1], np.arange(0, n, 64, dtype=np.uint64))
    assert np.array_equal(res["sa"], sa[::32].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::128].astype(np.uint64))


@pytest.mark.parametrize("bits,n,sigma", [(2, 70_001, 4), (3, 10_000, 5), (4, 33_333, 11), (8, 20_000, 256), (1, 4_097, 2), (2, 1, 4),
                                          (5, 64, 32), (7, 250_003, 100)])
@pytest.mark.parametrize("layout", ["le", "be"])
def test_random_compactstream_vs_oracle(eng, oracle, bits, n, sigma, layout):
    """inputtype=compactstream (K1 k_unpack_compact): container written by the oracle's numpy writer, symbols of every
    width incl. ones that straddle bytes and 64-bit words, words in either byte order; same result as the naive sort of
    the decoded symbols."""
    rng = np.random.default_rng(100 * bits + sigma)
    t = rng.integers(0, sigma, size=n, dtype=np.uint8)
    data = oracle.encode_compact(t, bits, layout)
    dec, b = oracle.decode_compact(data.tobytes())
    assert b == bits and np.array_equal(dec, t)
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    res, info = run(eng, data, "compactstream", preisarate=64, sasamplingrate=32, isasamplingrate=128)
    assert info["n"] == n
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["preisa"][:, 0], isa[::64].astype(np.uint64))
    assert np.array_equal(res["sa"], sa[::32].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::128].astype(np.uint64))
    # several blocks: same BWT
    if n > 1000:
        res3, _ = run(eng, data, "compactstream", numblocks=3, preisarate=64, sasamplingrate=32, isasamplingrate=128)
        assert np.array_equal(res3["bwt"], bwt) and np.array_equal(res3["sa"], res["sa"]) and np.array_equal(res3["isa"], res["isa"])


@pytest.mark.parametrize("seed,l", [(11, 8), (12, 1001), (13, 65536), (14, 250003), (15, 3), (16, 16), (17, 15)])
@pytest.mark.parametrize("itype", ["pac", "pacterm"])
def test_random_pac_vs_oracle(eng, oracle, seed, l, itype):
    rng = np.random.default_rng(seed)
    bases = rng.integers(0, 4, size=l, dtype=np.uint8)
    if itype == "pac" and l == 16:
        bases[0] = (bases[1] + 1) % 4  # keep the text primitive
    pac = oracle.encode_pac(bases)
    t = oracle.decode_pac(pac.tobytes(), term=(itype == "pacterm"))
    n = t.size
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    res, info = run(eng, pac, itype, preisarate=16, sasamplingrate=4, isasamplingrate=8)
    assert info["n"] == n
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["preisa"][:, 0], isa[::16].astype(np.uint64))
    assert np.array_equal(res["sa"], sa[::4].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::8].astype(np.uint64))


def test_repetitive_many_rounds(eng, oracle):
    """Mutated copies: LCPs in the hundreds force many prefix-doubling rounds."""
    rng = np.random.default_rng(21)
    base = rng.integers(0, 4, size=3000, dtype=np.uint8)
    parts = []
    for c in range(8):
        x = base.copy()
        pos = rng.integers(0, x.size, size=3)
        x[pos] = (x[pos] + 1) % 4
        parts.append(x)
    t = np.concatenate(parts)
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    res, info = run(eng, t, "bytestream", preisarate=64, sasamplingrate=32, isasamplingrate=64)
    assert info["sort_rounds"] >= 6
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["sa"], sa[::32].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::64].astype(np.uint64))


def test_checkbwt_property_8mbp(eng, oracle):
    """BASELINE config 1 size (8 Mbp ACGT bytestream, bwtonly=1): the restated checkbwt verifier
    (LF-walk against the circularly reversed text) accepts the GPU BWT + anchors."""
    rng = np.random.default_rng(1)
    t = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=8_000_000)]
    res, info = run(eng, t, "bytestream", bwtonly=True)
    assert info["preisarate"] == 64 and info["nsa"] == 0
    rc, checked = oracle.checkbwt(t, res["bwt"], res["preisa"], numthreads=8)
    assert rc == 1 and checked == t.size


def _check_all(oracle, res, t, prerate, sarate, isarate):
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["preisa"][:, 0], isa[::prerate].astype(np.uint64))
    assert np.array_equal(res["sa"], sa[::sarate].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::isarate].astype(np.uint64))


@pytest.mark.parametrize("sampling", ["auto", "walk"])
@pytest.mark.parametrize("itype", ["bytestream", "pacterm", "pac"])
def test_sampling_modes_agree(eng, oracle, itype, sampling):
    """Sampled SA/ISA straight from the suffix array (one block) == the reference's LF walk."""
    rng = np.random.default_rng(31)
    bases = rng.integers(0, 4, size=150_001, dtype=np.uint8)
    if itype == "bytestream":
        data, t = bases, bases
    else:
        data = oracle.encode_pac(bases)
        t = oracle.decode_pac(data.tobytes(), term=(itype == "pacterm"))
    res, info = run(eng, data, itype, preisarate=32, sasamplingrate=8, isasamplingrate=16, sampling=sampling)
    assert (info["walk_lf_steps"] > 0) == (sampling == "walk")
    _check_all(oracle, res, t, 32, 8, 16)


@pytest.mark.parametrize("sortpath", ["lsd", "auto"])
@pytest.mark.parametrize("n,sigma", [(500_000, 2), (400_000, 3)])
def test_many_small_groups_across_tiles(eng, oracle, n, sigma, sortpath):
    """Small alphabets: almost every suffix is tied after round 0, the runs of equal keys straddle
    the tiles of the resolve kernel (LSD sorter) / crowd the local digits of the finish kernel (MSD sorter)."""
    rng = np.random.default_rng(n + sigma)
    t = rng.integers(0, sigma, size=n, dtype=np.uint8)
    res, info = run(eng, t, "bytestream", preisarate=64, sasamplingrate=32, isasamplingrate=128, sortpath=sortpath)
    if sigma == 2 and sortpath == "lsd":
        assert info["sort_tied0"] > n // 2
    _check_all(oracle, res, t, 64, 32, 128)


@pytest.mark.parametrize("itype", ["bytestream", "pacterm"])
def test_long_runs_big_groups(eng, oracle, itype):
    """Runs of one symbol far longer than the first key: groups larger than the resolve kernel
    sorts in place, so the prefix-doubling rounds must finish the job."""
    rng = np.random.default_rng(77)
    parts = []
    for k in range(40):
        parts.append(np.full(int(rng.integers(50, 3000)), k % 4, dtype=np.uint8))
        parts.append(rng.integers(0, 4, size=int(rng.integers(1, 200)), dtype=np.uint8))
    bases = np.concatenate(parts)
    if itype == "bytestream":
        data, t = bases, bases
    else:
        data = oracle.encode_pac(bases)
        t = oracle.decode_pac(data.tobytes(), term=True)
    res, info = run(eng, data, itype, preisarate=16, sasamplingrate=4, isasamplingrate=8)
    assert info["sort_unresolved0"] > 0 and info["sort_rounds"] > 1
    _check_all(oracle, res, t, 16, 4, 8)


def test_tail_within_second_key(eng, oracle):
    """pacterm: suffixes that reach the terminator inside the second key compare by length."""
    for l in (17, 31, 33, 47, 48, 49, 63, 64, 65, 100):
        for fill in (0, 3):
            bases = np.full(l, fill, dtype=np.uint8)
            bases[l // 2] = (fill + 1) % 4
            data = oracle.encode_pac(bases)
            t = oracle.decode_pac(data.tobytes(), term=True)
            res, info = run(eng, data, "pacterm", preisarate=1, sasamplingrate=1, isasamplingrate=1)
            _check_all(oracle, res, t, 1, 1, 1)


def test_results_streamed_to_pinned_host(oracle):
    """b3m_build_params.host_sa / host_bwa: the sampled SA and BWA's packed BWT are delivered to a pinned host buffer during the build
    (chunks of the last sorting step) and equals the fetched one; also on the prefix-doubling path."""
    import torch
    from bwtb3m_b200 import Engine
    rng = np.random.default_rng(41)
    e = Engine(0)
    try:
        for kind in ("random", "repeats"):
            if kind == "random":
                bases = rng.integers(0, 4, size=1_500_003, dtype=np.uint8)
            else:
                u = rng.integers(0, 4, size=40_000, dtype=np.uint8)
                bases = np.concatenate([u, u[:30_000], rng.integers(0, 4, size=300_000, dtype=np.uint8), u])
            pac = oracle.encode_pac(bases)
            e.load_host(pac, "pacterm")
            n = bases.size + 1
            host = torch.full(((n + 31) // 32,), -1, dtype=torch.int64).pin_memory()
            hbwa = torch.full(((n - 1 + 15) // 16,), -1, dtype=torch.int32).pin_memory()
            e.build(sasamplingrate=32, isasamplingrate=64, host_sa_ptr=host.data_ptr(), host_bwa_ptr=hbwa.data_ptr())
            info = e.info()
            assert (info["sort_unresolved0"] > 0) == (kind == "repeats")
            res = e.fetch()
            assert np.array_equal(host.numpy().astype(np.uint64), res["sa"])
            words, primary, l2, seq_len = e.fetch_bwa()  # packed again into a fresh buffer
            if kind == "random":
                assert np.array_equal(hbwa.numpy().view(np.uint32), words)  # delivered during the build
            e.fetch_bwa(out_ptr=hbwa.data_ptr())  # no-op when delivered, a normal fetch otherwise
            assert np.array_equal(hbwa.numpy().view(np.uint32), words)
            t = oracle.decode_pac(pac.tobytes(), term=True)
            sa = oracle.sa_circular(t)
            assert np.array_equal(res["sa"], sa[::32].astype(np.uint64))
    finally:
        e.close()
