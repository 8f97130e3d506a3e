"""GPU: the multi-GPU build behind the C ABI (b3m_multi_*, `ngpus=` of bwtb3m; one process, one host thread
per GPU, peer stores) equals the single-GPU build bit for bit.  With one visible GPU only the ngpus=1 cases run
(the staging load path and the single strategy); the driver's multi-GPU boxes run all of them."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _single(data, itype, **kw):
    from bwtb3m_b200 import Engine
    e = Engine(0)
    try:
        e.load_host(data, itype)
        e.build(**kw)
        return e.info(), e.fetch()
    finally:
        e.close()


def _multi(ng, data, itype, **kw):
    from bwtb3m_b200 import MultiEngine
    m = MultiEngine(ng)
    try:
        m.load_host(data, itype)
        m.build(**kw)
        first = (m.engine.info(), m.engine.fetch(), m.stats())
        # a second build on the same handle (buffers are reused)
        m.load_host(data, itype)
        m.build(**kw)
        second = m.engine.fetch()
        for k in first[1]:
            assert np.array_equal(first[1][k], second[k]), "second build differs in " + k
        return first
    finally:
        m.close()


def _cases():
    from bwtb3m_b200 import workloads
    rng = np.random.default_rng(7)
    yield "pacterm", workloads.random_pac(2_000_003, 11), "xshard"
    yield "pac", workloads.random_pac(1_500_000, 12), "xshard"
    yield "bytestream", rng.integers(0, 256, size=700_001, dtype=np.uint8), "shard"
    yield "bytestream", np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=900_000)], "xshard"
    yield "pacterm", workloads.repetitive_pac(8, 40_000, 3), "single (text with long repeats)"


@pytest.mark.parametrize("ng", [1, 2, 4, 8])
def test_multi_equals_single(ng):
    if ng > _ngpus():
        pytest.skip("needs %d GPUs" % ng)
    for itype, data, want in _cases():
        for bwtonly in (False, True):
            kw = dict(sasamplingrate=32, isasamplingrate=1024, bwtonly=bwtonly)
            i1, r1 = _single(data, itype, **kw)
            kw["preisarate"] = i1["preisarate"]
            im, rm, st = _multi(ng, data, itype, **kw)
            assert st["strategy"] == ("single" if ng == 1 else want), (itype, st)
            assert im["n"] == i1["n"] and im["hist"] == i1["hist"]
            assert sorted(r1) == sorted(rm)
            for k in r1:
                assert np.array_equal(r1[k], rm[k]), "%s differs (%s, ngpus=%d, bwtonly=%s)" % (k, itype, ng, bwtonly)


def test_multi_forced_blocks_and_walk_take_the_block_path():
    from bwtb3m_b200 import workloads
    ng = min(_ngpus(), 2)
    data = workloads.random_pac(300_001, 5)
    i1, r1 = _single(data, "pacterm", numblocks=3, isasamplingrate=256)
    im, rm, st = _multi(ng, data, "pacterm", numblocks=3, isasamplingrate=256, preisarate=i1["preisarate"])
    assert st["strategy"] == "single" and im["numblocks"] == 3
    for k in r1:
        assert np.array_equal(r1[k], rm[k])


def test_cli_ngpus(tmp_path, oracle):
    """bwtb3m ngpus=N writes the files a one-GPU run writes; the reference's verifier accepts them."""
    from bwtb3m_b200 import workloads
    subprocess.check_call(["make", "-s", "-C", ROOT, "all"])
    ng = min(_ngpus(), 8)
    pac = workloads.random_pac(1_000_003, 21)
    fn = tmp_path / "t.pac"
    pac.tofile(fn)
    outs = {}
    for tag, n in (("one", 1), ("multi", ng)):
        out = tmp_path / (tag + ".bwt")
        r = subprocess.run([os.path.join(ROOT, "bin", "bwtb3m"), "inputtype=pacterm", "outputfilename=%s" % out, "ngpus=%d" % n, "verbose=1", str(fn)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        if n > 1:
            assert "%d GPUs" % n in r.stderr and "xshard" in r.stderr, r.stderr
        outs[tag] = {suf: open(str(out)[:-4] + suf, "rb").read() for suf in (".bwt", ".hist", ".sa", ".isa")}
    assert outs["one"] == outs["multi"]
    r = subprocess.run([os.path.join(ROOT, "bin", "bwtb3m"), "inputtype=pacterm", "ngpus=16", str(fn)], capture_output=True, text=True)  # a box has at most 8
    assert r.returncode != 0 and "does not exist" in r.stderr
