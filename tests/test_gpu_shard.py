"""GPU: suffix-range sharding (b3m_engine_shard_build / _finish, the multi-GPU fast path) on ONE
device: the key ranges are sorted one after the other into the same zeroed buffers (they write
disjoint places), and the result must equal the oracle exactly as a single-block build does."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sharded(eng, data, itype, nparts, **kw):
    import torch
    from bwtb3m_b200 import multigpu
    eng.load_host(data, itype)
    buf = multigpu.ShardBuffers(eng, kw.get("preisarate", 0), kw.get("sasamplingrate", 32), kw.get("isasamplingrate", 262144), False)
    unres = 0
    for part in range(nparts):
        unres += eng.shard_build(part, nparts, *buf.ptrs(), preisarate=buf.prerate, sasamplingrate=kw.get("sasamplingrate", 32),
                                 isasamplingrate=kw.get("isasamplingrate", 262144))
    torch.cuda.synchronize()
    if unres:
        return None, unres
    eng.shard_finish(nparts, *buf.ptrs())
    return eng.fetch(), 0


@pytest.fixture(scope="module")
def eng():
    from bwtb3m_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("nparts", [1, 2, 3, 8])
@pytest.mark.parametrize("itype,n", [("pacterm", 200_003), ("pac", 150_000), ("bytestream", 120_001), ("pacterm", 17)])
def test_shards_equal_oracle(eng, oracle, itype, n, nparts):
    rng = np.random.default_rng(n + nparts)
    if itype == "bytestream":
        data = rng.integers(0, 256, size=n, dtype=np.uint8)
        t = data
    else:
        bases = rng.integers(0, 4, size=n, dtype=np.uint8)
        data = oracle.encode_pac(bases)
        t = oracle.decode_pac(data.tobytes(), term=(itype == "pacterm"))
    res, unres = sharded(eng, data, itype, nparts, preisarate=64, sasamplingrate=8, isasamplingrate=32)
    assert unres == 0
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    assert np.array_equal(res["bwt"], bwt)
    assert np.array_equal(res["preisa"][:, 0], isa[::64].astype(np.uint64))
    assert np.array_equal(res["sa"], sa[::8].astype(np.uint64))
    assert np.array_equal(res["isa"], isa[::32].astype(np.uint64))
    if itype == "pacterm":
        words, primary, l2, seq_len = eng.fetch_bwa()
        assert primary == int(isa[0]) and seq_len == t.size - 1


def test_skewed_text_balances_and_matches(eng, oracle):
    """Low-entropy DNA (long A runs between random stretches): ranges are cut at bin boundaries of the
    key histogram; either the shards finish and match, or they report unresolved suffixes."""
    rng = np.random.default_rng(5)
    parts = []
    for k in range(200):
        parts.append(np.zeros(int(rng.integers(1, 40)), dtype=np.uint8))
        parts.append(rng.integers(0, 4, size=int(rng.integers(50, 400)), dtype=np.uint8))
    bases = np.concatenate(parts)
    data = oracle.encode_pac(bases)
    t = oracle.decode_pac(data.tobytes(), term=True)
    res, unres = sharded(eng, data, "pacterm", 4, preisarate=16, sasamplingrate=4, isasamplingrate=8)
    if unres == 0:
        sa = oracle.sa_circular(t)
        bwt, isa = oracle.bwt_from_sa(t, sa)
        assert np.array_equal(res["bwt"], bwt)
        assert np.array_equal(res["sa"], sa[::4].astype(np.uint64))
        assert np.array_equal(res["isa"], isa[::8].astype(np.uint64))


def test_repeats_are_reported_unresolved(eng, oracle):
    """Two copies of the same 5 kbp sequence: ties beyond both sort keys -> the shard path must say so."""
    rng = np.random.default_rng(6)
    u = rng.integers(0, 4, size=5000, dtype=np.uint8)
    bases = np.concatenate([u, rng.integers(0, 4, size=100, dtype=np.uint8), u])
    data = oracle.encode_pac(bases)
    res, unres = sharded(eng, data, "pacterm", 2)
    assert res is None and unres > 0
