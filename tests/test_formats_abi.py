"""CPU-only: the C-ABI library loads and exports every symbol include/b3m.h declares, the file
formats round-trip, the CLIs keep the reference's argument/error conventions, and the library
fails loudly without a CUDA device (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", ROOT, "all"])


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    from bwtb3m_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b3m.h")).read()
    declared = set(re.findall(r"\b(b3m_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libb3m.so does not export " + name
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert b"sm_100a" in L.b3m_version()


def test_parse_inputtype():
    from bwtb3m_b200 import _lib
    L = _lib.lib()
    assert [L.b3m_parse_inputtype(x) for x in (b"bytestream", b"compactstream", b"pac", b"pacterm")] == [0, 1, 2, 3]
    assert L.b3m_parse_inputtype(b"lz4") == -1 and L.b3m_parse_inputtype(b"utf-8") == -1


@pytest.mark.parametrize("kind", ["dna", "bytes", "runs", "single", "long_runs", "one_symbol"])
def test_rl_container_roundtrip(tmp_path, kind):
    from bwtb3m_b200 import files
    rng = np.random.default_rng(7)
    if kind == "dna":
        s = rng.integers(1, 5, size=300_000, dtype=np.uint8)
        s[1234] = 0
    elif kind == "bytes":
        s = rng.integers(0, 256, size=100_000, dtype=np.uint8)
    elif kind == "runs":
        s = np.repeat(rng.integers(0, 4, size=20_000, dtype=np.uint8), rng.integers(1, 40, size=20_000))
    elif kind == "single":
        s = np.array([65], dtype=np.uint8)
    elif kind == "long_runs":
        s = np.repeat(np.array([3, 1, 3, 0, 2], dtype=np.uint8), [70_000, 255, 256, 1, 100_000])
    else:
        s = np.full(50_000, 7, dtype=np.uint8)
    fn = str(tmp_path / "x.bwt")
    files.write_bwt_host(fn, s)
    assert files.bwt_length(fn) == s.size
    assert np.array_equal(files.read_bwt(fn, numthreads=3), s)
    # the streaming reader of bwtb3mdecoderl gives the same bytes
    out = subprocess.run([os.path.join(BIN, "bwtb3mdecoderl"), fn], capture_output=True, check=True).stdout
    assert out == s.tobytes()
    # several files are addressed as one sequence (RLDecoder(vector<string>,...), bwtb3mdecoderl.cpp:27)
    out2 = subprocess.run([os.path.join(BIN, "bwtb3mdecoderl"), fn, fn], capture_output=True, check=True).stdout
    assert out2 == s.tobytes() * 2


def test_block_sym_histograms_and_rank(tmp_path):
    """getBlockSymHistograms / SparseRank::rankm of the reference's random-access consumer
    (/root/reference/src/bwtdecodeblock.cpp:210-242,356-365): the `.sparserank` file holds, per block of the container,
    the occurrences of every symbol before the block (big-endian uint64), and rank queries from it equal a prefix count."""
    from bwtb3m_b200 import files
    rng = np.random.default_rng(12)
    # runs of uneven length over symbols 1..4 plus one 0, long enough for several blocks of 4096 runs
    lens = rng.integers(1, 9, size=30_000)
    syms = rng.integers(1, 5, size=lens.size).astype(np.uint8)
    L = np.repeat(syms, lens)
    L[L.size // 3] = 0
    fn = str(tmp_path / "x.bwt")
    files.write_bwt_host(fn, L)
    nb = files.block_sym_histograms(fn, fn + ".sparserank", 0, 4, numthreads=3)
    raw = np.fromfile(fn + ".sparserank", dtype=">u8").reshape(nb, 5)
    assert nb >= 2 and np.all(raw[0] == 0) and np.all(np.diff(raw.astype(np.int64), axis=0) >= 0)
    assert raw[-1].sum() < L.size  # counts BEFORE the last block
    cum = np.zeros((5, L.size + 1), dtype=np.int64)
    for c in range(5):
        cum[c, 1:] = np.cumsum(L == c)
    for i in [0, 1, 2, L.size // 3, L.size // 3 + 1, L.size - 1, L.size] + [int(x) for x in rng.integers(0, L.size, size=40)]:
        for c in range(5):
            assert files.bwt_rank(fn, fn + ".sparserank", 0, 4, c, i) == cum[c, i], (c, i)
    with pytest.raises(Exception):
        files.block_sym_histograms(fn, fn + ".bad", 1, 4)  # symbol 0 occurs: outside the stated range


def test_to_bwa_host_tool_against_oracle(tmp_path, oracle):
    """bwtb3mtobwa is a host tool (files in, files out, no GPU): .bwt + .sa of a pacterm text -> BWA's .bwt / .sa,
    byte for byte what the restated MausFmToBwaConversion gives; texts without exactly one terminator are refused."""
    from bwtb3m_b200 import files
    rng = np.random.default_rng(21)
    for l, sarate in ((5, 1), (1000, 4), (70_001, 32)):
        t = np.concatenate([rng.integers(1, 5, size=l, dtype=np.uint8), np.zeros(1, dtype=np.uint8)])
        sa = oracle.sa_circular(t)
        bwt, isa = oracle.bwt_from_sa(t, sa)
        fn = str(tmp_path / ("t%d.bwt" % l))
        files.write_bwt_host(fn, bwt)
        samples = sa[::sarate].astype(np.uint64)
        with open(fn[:-4] + ".sa", "wb") as f:
            f.write(np.array([sarate, samples.size], dtype=np.uint64).tobytes())
            f.write(samples.tobytes())
        files.to_bwa(fn, fn + ".bwa", fn + ".bwasa")
        ob, osa = oracle.to_bwa(bwt, samples, sarate)
        assert open(fn + ".bwa", "rb").read() == ob
        assert open(fn + ".bwasa", "rb").read() == osa
    bad = str(tmp_path / "bad.bwt")
    files.write_bwt_host(bad, rng.integers(1, 5, size=100, dtype=np.uint8))  # no terminator
    open(bad[:-4] + ".sa", "wb").write(np.array([4, 25] + [0] * 25, dtype=np.uint64).tobytes())
    with pytest.raises(Exception):
        files.to_bwa(bad, bad + ".bwa", bad + ".bwasa")


def test_rl_container_rejects_garbage(tmp_path):
    from bwtb3m_b200 import files
    from bwtb3m_b200.engine import B3MError
    fn = str(tmp_path / "bad.bwt")
    open(fn, "wb").write(b"not a container at all, but long enough to hold a header........")
    with pytest.raises(B3MError):
        files.bwt_length(fn)
    with pytest.raises(B3MError):
        files.read_bwt(str(tmp_path / "missing.bwt"))


def test_cli_help_and_errors(tmp_path):
    exe = os.path.join(BIN, "bwtb3m")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1  # help is delivered as an error, /root/reference/src/bwtb3m.cpp:32-59,67-71
    for key in ("inputtype=", "outputfilename=", "sasamplingrate=[32]", "isasamplingrate=[262144]", "mem=", "numthreads=", "bwtonly=[0]",
                "tmpprefix=", "sparsetmpprefix=", "copyinputtomemory=", "largelcpthres=[16384]", "verbose=[0]"):
        assert key in r.stderr, key
    assert subprocess.run([exe, "-h"], capture_output=True).returncode == 1
    r = subprocess.run([exe, "inputtype=lz4", str(tmp_path / "nofile")], capture_output=True, text=True)
    assert r.returncode == 1 and "lz4" in r.stderr
    inp = tmp_path / "in.txt"
    inp.write_bytes(b"ACGTACGTTGCA")
    r = subprocess.run([exe, "outputfilename=" + str(tmp_path / "o.bwt"), str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode == 1
    for tool in ("bwtb3mtobwa", "bwtcomputessa", "bwtb3mdecoderl"):
        assert subprocess.run([os.path.join(BIN, tool)], capture_output=True).returncode == 1


@pytest.mark.skipif(has_gpu(), reason="checks the behaviour on a box without a GPU")
def test_no_cpu_fallback(tmp_path):
    """Without a CUDA device every compute entry point fails loudly."""
    from bwtb3m_b200 import Engine, B3MError, files
    with pytest.raises(B3MError, match="no CPU fallback"):
        Engine(0)
    inp = tmp_path / "in.txt"
    inp.write_bytes(b"ACGTACGTTGCA")
    with pytest.raises(B3MError, match="no CPU fallback"):
        files.compute_bwt(str(inp), outputfilename=str(tmp_path / "o.bwt"))
    r = subprocess.run([os.path.join(BIN, "bwtb3m"), "outputfilename=" + str(tmp_path / "o.bwt"), str(inp)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert not (tmp_path / "o.bwt").exists()


def test_product_package_never_uses_oracle():
    """The oracle is test infrastructure: nothing in the product package or the CLIs may import,
    link, load or call it."""
    bad = re.compile(r"import\s+oracle|from\s+oracle|libb3m_oracle|\borc_[a-z]|oracle/.*\.so|-lb3m_oracle")
    roots = [os.path.join(ROOT, "bwtb3m_b200"), os.path.join(ROOT, "cli"), os.path.join(ROOT, "include")]
    for root in roots:
        for dp, _, fns in os.walk(root):
            for f in fns:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    src = open(os.path.join(dp, f), errors="replace").read()
                    assert not bad.search(src), os.path.join(dp, f)
    mk = open(os.path.join(ROOT, "Makefile")).read()
    assert "b3m_oracle" not in mk.replace("$(MAKE) -s -C oracle", "")


def test_sampled_and_preisa_layouts(tmp_path):
    """Layouts the reference's tools read: .sa/.isa = [rate][count][values] native uint64
    (sasubsample.cpp:34-58); .preisa = (rank,pos) pairs, size % 16 == 0 (hwtPreIsaToIsa.cpp:55-77)."""
    from bwtb3m_b200 import files
    fn = str(tmp_path / "x.sa")
    np.array([32, 3, 10, 20, 30], dtype=np.uint64).tofile(fn)
    rate, v = files.read_sampled(fn)
    assert rate == 32 and v.tolist() == [10, 20, 30]
    fn = str(tmp_path / "x.preisa")
    np.array([[5, 0], [9, 64]], dtype=np.uint64).tofile(fn)
    assert files.read_preisa(fn).tolist() == [[5, 0], [9, 64]]
    assert os.path.getsize(fn) % 16 == 0


def test_sasubsample_tool(tmp_path):
    """cli/sasubsample: [rate][count][values] -> every s-th value, rate*s (reference: src/sasubsample.cpp:34-58)."""
    import subprocess
    import numpy as np
    exe = os.path.join(ROOT, "bin", "sasubsample")
    if not os.path.exists(exe):
        pytest.skip("bin/sasubsample not built")
    vals = np.arange(100, 100 + 37, dtype=np.uint64)
    src = np.concatenate([np.array([32, vals.size], dtype=np.uint64), vals]).tobytes()
    out = subprocess.run([exe, "-s", "4"], input=src, capture_output=True, check=True).stdout
    a = np.frombuffer(out, dtype=np.uint64)
    assert a[0] == 128 and a[1] == (37 + 3) // 4 and np.array_equal(a[2:], vals[::4])
    assert subprocess.run([exe, "-s", "3"], input=src, capture_output=True).returncode != 0
