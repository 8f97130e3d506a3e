"""Pins the CPU oracle against the known-answer vectors of SURVEY.md section 4 (derived from
the reference's own BWT/ISA definition, /root/reference/src/lcpbit.cpp:3658-3669,3688-3712, text
"abbab#" from lcpbit.cpp:4053) and against its independent naive rotation sort, in the pattern
of the reference's exhaustive self-test (lcpbit.cpp:3777-3794,4054,4065-4066)."""
import itertools

import numpy as np
import pytest

KATS = [
    (b"abbab#", [5, 3, 0, 4, 2, 1], b"bb#aba", [2, 5, 4, 1, 3, 0]),
    (b"banana", [5, 3, 1, 0, 4, 2], b"nnbaaa", [3, 2, 5, 1, 4, 0]),
    (b"mississippi", [10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2], b"pssmipissii", [4, 3, 10, 8, 2, 9, 7, 1, 6, 5, 0]),
    (b"ACGTACGTTGCA", [11, 0, 4, 10, 1, 5, 9, 2, 6, 3, 8, 7], b"CATGAATCCGTG", [1, 4, 7, 9, 2, 5, 8, 11, 10, 6, 3, 0]),
]


def is_primitive(t):
    n = len(t)
    for p in range(1, n):
        if n % p == 0 and all(t[i] == t[i % p] for i in range(n)):
            return False
    return True


def all_pairs(bwt_isa, n, rate):
    isa = bwt_isa
    return np.array([[isa[p], p] for p in range(0, n, rate)], dtype=np.uint64)


@pytest.mark.parametrize("text,sa,bwt,isa", KATS)
def test_kat_naive_and_fast(oracle, text, sa, bwt, isa):
    t = np.frombuffer(text, dtype=np.uint8)
    for sorter in (oracle.naive_sa, oracle.sa_circular):
        s = sorter(t)
        assert s.tolist() == sa
        b, i = oracle.bwt_from_sa(t, s)
        assert b.tobytes() == bwt
        assert i.tolist() == isa


def test_kat_pacterm(oracle):
    # .pac bytes of ACGTTGCA are 1b e4 00 00; pacterm symbols 1 2 3 4 4 3 2 1 0
    pac = bytes([0x1B, 0xE4, 0x00, 0x00])
    assert oracle.encode_pac([0, 1, 2, 3, 3, 2, 1, 0]).tobytes() == pac
    t = oracle.decode_pac(pac, term=True)
    assert t.tolist() == [1, 2, 3, 4, 4, 3, 2, 1, 0]
    assert oracle.decode_pac(pac, term=False).tolist() == [0, 1, 2, 3, 3, 2, 1, 0]
    sa = oracle.sa_circular(t)
    assert sa.tolist() == [8, 7, 0, 6, 1, 5, 2, 4, 3]
    b, isa = oracle.bwt_from_sa(t, sa)
    assert b.tolist() == [1, 2, 0, 3, 1, 4, 2, 4, 3]
    assert isa[0] == 2  # BWA primary


@pytest.mark.parametrize("text,sa,bwt,isa", KATS)
@pytest.mark.parametrize("nblocks", [1, 2, 3, 5])
def test_kat_b3m_blocks(oracle, text, sa, bwt, isa, nblocks):
    t = np.frombuffer(text, dtype=np.uint8)
    b, pp, _ = oracle.b3m(t, nblocks=nblocks, rate=2, nthreads=2)
    assert b.tobytes() == bwt
    for r, p in pp.tolist():
        assert isa[p] == r
    assert [p for _, p in pp.tolist()] == list(range(0, len(text), 2))
    rc, checked = oracle.checkbwt(t, b, pp, numthreads=3)
    assert rc == 1 and checked == len(text)
    s32, i32 = oracle.ssa(b, pp, sarate=2, isarate=4, nthreads=2)
    assert s32.tolist() == [sa[r] for r in range(0, len(text), 2)]
    assert i32.tolist() == [isa[p] for p in range(0, len(text), 4)]


def _exhaustive(alpha, length):
    for tup in itertools.product(range(alpha), repeat=length):
        yield np.array(list(tup) + [alpha], dtype=np.uint8)  # + unique '#'-like last symbol as in lcpbit test


@pytest.mark.parametrize("alpha,length", [(2, 6), (3, 5), (4, 5)])
def test_exhaustive_small_alphabets(oracle, alpha, length):
    """All strings of the given length (+ a final separator symbol) in the spirit of the
    reference's exhaustive lcpbit self-test: fast sorter and every block count agree with the
    naive rotation sort."""
    for t in _exhaustive(alpha, length):
        sa = oracle.naive_sa(t)
        assert oracle.sa_circular(t).tolist() == sa.tolist()
        bwt, isa = oracle.bwt_from_sa(t, sa)
        for nb in (2, 3, 4):
            b, pp, _ = oracle.b3m(t, nblocks=nb, rate=1, nthreads=1)
            assert b.tolist() == bwt.tolist(), (t.tolist(), nb)
            assert pp[:, 0].tolist() == isa.tolist(), (t.tolist(), nb)


def test_exhaustive_no_separator_primitive(oracle):
    """Terminator-free circular texts (bytestream semantics): all primitive strings of length 8
    over 2 symbols and of length 6 over 3 symbols."""
    for alpha, length in ((2, 8), (3, 6)):
        for tup in itertools.product(range(alpha), repeat=length):
            if not is_primitive(tup):
                continue
            t = np.array(tup, dtype=np.uint8)
            sa = oracle.naive_sa(t)
            assert oracle.sa_circular(t).tolist() == sa.tolist()
            bwt, isa = oracle.bwt_from_sa(t, sa)
            for nb in (2, 3):
                b, pp, _ = oracle.b3m(t, nblocks=nb, rate=1, nthreads=1)
                assert b.tolist() == bwt.tolist(), (tup, nb)
                assert pp[:, 0].tolist() == isa.tolist(), (tup, nb)


@pytest.mark.parametrize("seed,n,sigma", [(1, 1000, 4), (2, 4097, 4), (3, 3000, 256), (4, 2500, 2)])
def test_random_blocks_vs_naive(oracle, seed, n, sigma):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, sigma, size=n, dtype=np.uint8)
    sa = oracle.naive_sa(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    assert oracle.sa_circular(t).tolist() == sa.tolist()
    for nb in (1, 2, 7):
        b, pp, st = oracle.b3m(t, nblocks=nb, rate=64, nthreads=4)
        assert np.array_equal(b, bwt)
        assert np.array_equal(pp[:, 0], isa[::64].astype(np.uint64))
        rc, checked = oracle.checkbwt(t, b, pp, numthreads=8)
        assert rc == 1 and checked == n
        s, i = oracle.ssa(b, pp, sarate=32, isarate=128, nthreads=4)
        assert np.array_equal(s, sa[::32].astype(np.uint64))
        assert np.array_equal(i, isa[::128].astype(np.uint64))


def test_repetitive_large_lcp(oracle):
    """Mutated copies: look-ahead (lcpnext) far above the block-local scale, and above a small
    largelcpthres so that the exact (escape) LCP computation is exercised (A4)."""
    rng = np.random.default_rng(7)
    base = rng.integers(0, 4, size=700, dtype=np.uint8)
    parts = []
    for c in range(6):
        x = base.copy()
        pos = rng.integers(0, x.size, size=2)
        x[pos] = (x[pos] + 1 + rng.integers(0, 3, size=2)) % 4
        parts.append(x)
    t = np.concatenate(parts)
    sa = oracle.naive_sa(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    for nb in (2, 6, 9):
        b, pp, st = oracle.b3m(t, nblocks=nb, rate=16, largelcpthres=8, nthreads=4)
        assert np.array_equal(b, bwt)
        if nb != 9:  # block boundaries on copy boundaries: look-ahead spans a whole copy prefix
            assert st["max_lcpnext"] > 8
        assert np.array_equal(pp[:, 0], isa[::16].astype(np.uint64))


def test_checkbwt_detects_corruption(oracle):
    rng = np.random.default_rng(11)
    t = rng.integers(0, 4, size=5000, dtype=np.uint8)
    b, pp, _ = oracle.b3m(t, nblocks=3, rate=64, nthreads=2)
    assert oracle.checkbwt(t, b, pp, numthreads=4)[0] == 1
    bad = b.copy()
    i, j = 100, 4000
    while bad[i] == bad[j]:
        j += 1
    bad[i], bad[j] = bad[j], bad[i]
    assert oracle.checkbwt(t, bad, pp, numthreads=4)[0] == 0


def test_bwa_export_kat(oracle):
    """BWA .bwt/.sa layout on the pacterm KAT: primary = 2."""
    t = np.array([1, 2, 3, 4, 4, 3, 2, 1, 0], dtype=np.uint8)
    b, pp, _ = oracle.b3m(t, nblocks=2, rate=1, nthreads=1)
    assert b.tolist() == [1, 2, 0, 3, 1, 4, 2, 4, 3]
    sa, isa = oracle.ssa(b, pp, sarate=2, isarate=1, nthreads=1)
    assert sa.tolist() == [8, 0, 1, 2, 3]
    bwt_bytes, sa_bytes = oracle.to_bwa(b, sa, 2)
    w = np.frombuffer(bwt_bytes[:40], dtype=np.uint64)
    assert w[0] == 2                       # primary = ISA[0]
    assert w[1:].tolist() == [2, 4, 6, 8]  # L2[1..4]
    word = np.frombuffer(bwt_bytes[40:], dtype=np.uint32)
    assert word.size == 1
    syms = [(int(word[0]) >> ((15 - i) * 2)) & 3 for i in range(8)]
    assert syms == [0, 1, 2, 0, 3, 1, 3, 2]  # BWT without the terminator row, A=0
    h = np.frombuffer(sa_bytes, dtype=np.uint64)
    assert h[0] == 2 and h[5] == 2 and h[6] == 8
    assert h[7:].tolist() == [0, 1, 2, 3]   # SA[2],SA[4],SA[6],SA[8]; SA[0] implicit


def test_default_numblocks_formula(oracle):
    # mem=2 GiB, 8 threads: tblock = 0.95*2^31/40 = 51,002,736 -> 48,000,001 symbols: ceil(fs/threads)=6,000,001 wins
    assert oracle.default_numblocks(48_000_001, 2 << 30, 8) == 8
    assert oracle.default_numblocks(3_100_000_001, 2 << 30, 8) == 61
