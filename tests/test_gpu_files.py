"""GPU parity, file level: bwtb3m -> .bwt/.hist/.preisa/.sa/.isa, bwtcomputessa, bwtb3mtobwa
through the C ABI and the command line tools, checked against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def make_pac(oracle, l, seed):
    rng = np.random.default_rng(seed)
    bases = rng.integers(0, 4, size=l, dtype=np.uint8)
    pac = oracle.encode_pac(bases)
    return pac, oracle.decode_pac(pac.tobytes(), term=True)


@pytest.mark.parametrize("numblocks", [0, 3])
def test_compute_bwt_files_pacterm(tmp_path, oracle, numblocks):
    from bwtb3m_b200 import files
    pac, t = make_pac(oracle, 100_003, 3)
    fn = tmp_path / "g.pac"
    pac.tofile(fn)
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    res = files.compute_bwt(str(fn), inputtype="pacterm", outputfilename=str(tmp_path / "g.bwt"), sasamplingrate=32,
                            isasamplingrate=256, numblocks=numblocks)
    assert res["n"] == t.size and res["numblocks"] == max(numblocks, 1)
    assert res["bwtfn"] == str(tmp_path / "g.bwt") and res["safn"].endswith("g.sa") and res["isafn"].endswith("g.isa")
    assert files.bwt_length(res["bwtfn"]) == t.size
    got = files.read_bwt(res["bwtfn"])
    assert np.array_equal(got, bwt)
    # the device encoder (K8) writes the same bytes as the host encoder of the container
    files.write_bwt_host(str(tmp_path / "h.bwt"), bwt)
    assert open(res["bwtfn"], "rb").read() == open(tmp_path / "h.bwt", "rb").read()
    rate, v = files.read_sampled(res["safn"])
    assert rate == 32 and np.array_equal(v, sa[::32].astype(np.uint64))
    rate, v = files.read_sampled(res["isafn"])
    assert rate == 256 and np.array_equal(v, isa[::256].astype(np.uint64))
    hist = files.read_hist(res["histfn"])
    assert hist == {int(s): int(c) for s, c in zip(*np.unique(t, return_counts=True))}
    assert not os.path.exists(tmp_path / "g.preisa")  # removed in bwtonly=0 mode (reference ChangeLog 0.0.49)
    # BWA export
    files.to_bwa(res["bwtfn"], str(tmp_path / "bwa.bwt"), str(tmp_path / "bwa.sa"))
    ob, osa = oracle.to_bwa(bwt, sa[::32].astype(np.uint64), 32)
    assert open(tmp_path / "bwa.bwt", "rb").read() == ob
    assert open(tmp_path / "bwa.sa", "rb").read() == osa


def test_bwtonly_then_computessa(tmp_path, oracle):
    """README staging: bwtonly=1 leaves .bwt + .preisa (+.meta); bwtcomputessa resumes from them."""
    from bwtb3m_b200 import files
    rng = np.random.default_rng(5)
    t = rng.integers(0, 256, size=60_001, dtype=np.uint8)
    fn = tmp_path / "t.bin"
    t.tofile(fn)
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    res = files.compute_bwt(str(fn), outputfilename=str(tmp_path / "t.bwt"), bwtonly=True, numblocks=2)
    assert res["safn"] == "" and res["isafn"] == ""
    assert np.array_equal(files.read_bwt(res["bwtfn"]), bwt)
    pre = files.read_preisa(res["preisafn"])
    assert np.array_equal(pre[:, 1], np.arange(0, t.size, 64, dtype=np.uint64))  # preisa rate 64 when bwtonly=1 (ChangeLog:281)
    assert np.array_equal(pre[:, 0], isa[::64].astype(np.uint64))
    assert open(res["metafn"], "rb").read() == (64).to_bytes(8, "big")
    rc, checked = oracle.checkbwt(t, bwt, pre, numthreads=4)
    assert rc == 1 and checked == t.size
    files.compute_ssa(res["bwtfn"], sasamplingrate=16, isasamplingrate=128)
    rate, v = files.read_sampled(str(tmp_path / "t.sa"))
    assert rate == 16 and np.array_equal(v, sa[::16].astype(np.uint64))
    rate, v = files.read_sampled(str(tmp_path / "t.isa"))
    assert rate == 128 and np.array_equal(v, isa[::128].astype(np.uint64))
    # ref_sa / ref_isa comparison arguments of the reference tool
    files.compute_ssa(res["bwtfn"], sasamplingrate=16, isasamplingrate=128, ref_sa=str(tmp_path / "t.sa"), ref_isa=str(tmp_path / "t.isa"))
    np.array([16, 1, 0], dtype=np.uint64).tofile(tmp_path / "wrong.sa")
    from bwtb3m_b200 import B3MError
    with pytest.raises(B3MError):
        files.compute_ssa(res["bwtfn"], sasamplingrate=16, isasamplingrate=128, ref_sa=str(tmp_path / "wrong.sa"))


def test_computessa_pacterm_irregular_anchors(tmp_path, oracle):
    """The .preisa of the reference has arbitrary order and spacing (sortPreIsa.cpp:103-127)."""
    from bwtb3m_b200 import files
    pac, t = make_pac(oracle, 50_000, 9)
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    files.write_bwt_host(str(tmp_path / "p.bwt"), bwt)
    rng = np.random.default_rng(1)
    pos = np.unique(rng.integers(0, t.size, size=37))
    rng.shuffle(pos)
    np.stack([isa[pos].astype(np.uint64), pos.astype(np.uint64)], axis=1).tofile(tmp_path / "p.preisa")
    files.compute_ssa(str(tmp_path / "p.bwt"), sasamplingrate=8, isasamplingrate=8)
    assert np.array_equal(files.read_sampled(str(tmp_path / "p.sa"))[1], sa[::8].astype(np.uint64))
    assert np.array_equal(files.read_sampled(str(tmp_path / "p.isa"))[1], isa[::8].astype(np.uint64))


def test_cli_pipeline(tmp_path, oracle):
    """README pipeline with the command line tools: bwtb3m -> bwtb3mdecoderl, bwtb3mtobwa."""
    pac, t = make_pac(oracle, 30_000, 11)
    fn = tmp_path / "r.pac"
    pac.tofile(fn)
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    r = subprocess.run([os.path.join(BIN, "bwtb3m"), "inputtype=pacterm", "outputfilename=" + str(tmp_path / "r.bwt"), "sasamplingrate=32",
                        "isasamplingrate=64", "mem=512k", "verbose=1", str(fn)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "[M]" in r.stderr and "runtime" in r.stderr
    assert "numblocks=2" in r.stderr  # mem=512k bounds the block size: 512 KiB / 30 B = 17476 suffixes per block
    out = subprocess.run([os.path.join(BIN, "bwtb3mdecoderl"), str(tmp_path / "r.bwt")], capture_output=True, check=True).stdout
    assert out == bwt.tobytes()
    r = subprocess.run([os.path.join(BIN, "bwtb3mtobwa"), str(tmp_path / "r.bwt"), str(tmp_path / "bwa.bwt"), str(tmp_path / "bwa.sa")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ob, osa = oracle.to_bwa(bwt, sa[::32].astype(np.uint64), 32)
    assert open(tmp_path / "bwa.bwt", "rb").read() == ob and open(tmp_path / "bwa.sa", "rb").read() == osa
    primary = int(np.frombuffer(ob[:8], dtype=np.uint64)[0])
    assert primary == isa[0]


def test_fetch_runs(oracle):
    from bwtb3m_b200 import Engine
    import ctypes as C
    rng = np.random.default_rng(2)
    t = np.repeat(rng.integers(0, 3, size=3000, dtype=np.uint8), rng.integers(1, 9, size=3000))
    t[0] = 3
    sa = oracle.sa_circular(t)
    bwt, _ = oracle.bwt_from_sa(t, sa)
    e = Engine(0)
    e.load_host(t, "bytestream")
    e.build(bwtonly=True)
    syms, lens = e.fetch_runs()
    assert np.array_equal(np.repeat(syms, lens.astype(np.int64)), bwt)
    assert np.all(syms[1:] != syms[:-1])
    e.close()


@pytest.mark.parametrize("l", [1, 15, 16, 17, 1000, 65537])
def test_engine_fetch_bwa_matches_oracle(oracle, l):
    """K9: the device-packed BWA words, primary and L2 equal the oracle's BWA export of the same BWT."""
    from bwtb3m_b200 import Engine
    rng = np.random.default_rng(900 + l)
    bases = rng.integers(0, 4, size=l, dtype=np.uint8)
    pac = oracle.encode_pac(bases)
    eng = Engine(0)
    try:
        eng.load_host(pac, "pacterm")
        eng.build(sasamplingrate=4, isasamplingrate=8)
        res = eng.fetch()
        words, primary, l2, seq_len = eng.fetch_bwa()
    finally:
        eng.close()
    ob, osa = oracle.to_bwa(res["bwt"], res["sa"], 4)
    assert seq_len == l
    assert primary == int(np.frombuffer(ob[:8], dtype=np.uint64)[0])
    assert l2[1:] == np.frombuffer(ob[8:40], dtype=np.uint64).tolist()
    assert words.tobytes() == ob[40:40 + 4 * words.size]
    # BWA's .sa payload is the rank-sampled SA without its first entry
    assert res["sa"][1:].tobytes() == osa[56:56 + 8 * (res["sa"].size - 1)]


@pytest.mark.parametrize("itype", ["pacterm", "bytestream"])
def test_checkbwt_tool_and_abi(tmp_path, oracle, itype):
    """checkbwt on the GPU (b3m_check_bwt + cli/checkbwt): accepts the files bwtb3m wrote, rejects a
    BWT with two rows swapped (same symbol counts) and one with a changed symbol."""
    from bwtb3m_b200 import files
    if itype == "pacterm":
        data, t = make_pac(oracle, 120_003, 9)
        fn = tmp_path / "g.pac"
    else:
        t = np.random.default_rng(10).integers(0, 200, size=90_001, dtype=np.uint8)
        data = t
        fn = tmp_path / "t.bin"
    data.tofile(fn)
    res = files.compute_bwt(str(fn), inputtype=itype, outputfilename=str(tmp_path / "x.bwt"), bwtonly=True)
    ok, bad = files.check_bwt(res["bwtfn"], str(fn), inputtype=itype)
    assert ok and bad == 0
    r = subprocess.run([os.path.join(BIN, "checkbwt"), "-i", itype, res["bwtfn"], str(fn)], capture_output=True, text=True)
    assert r.returncode == 0 and "[V] gok=1" in r.stderr
    bwt = files.read_bwt(res["bwtfn"])
    # swap two neighbouring rows that hold different symbols: histogram unchanged, LF walk derails
    k = int(np.nonzero(bwt[1000:-1] != bwt[1001:])[0][0]) + 1000
    broken = bwt.copy()
    broken[k], broken[k + 1] = bwt[k + 1], bwt[k]
    files.write_bwt_host(res["bwtfn"], broken)
    ok, bad = files.check_bwt(res["bwtfn"], str(fn), inputtype=itype)
    assert not ok and bad > 0
    r = subprocess.run([os.path.join(BIN, "checkbwt"), "-i", itype, res["bwtfn"], str(fn)], capture_output=True, text=True)
    assert r.returncode == 0 and "[V] gok=0" in r.stderr
    # a changed symbol changes the counts
    broken = bwt.copy()
    j = 5000 if bwt[5000] != 0 else 5001  # not the terminator row
    broken[j] = (bwt[j] % 4) + 1 if itype == "pacterm" else (int(bwt[j]) + 1) % 200
    files.write_bwt_host(res["bwtfn"], broken)
    ok, bad = files.check_bwt(res["bwtfn"], str(fn), inputtype=itype)
    assert not ok


def test_fasta_to_compactstream_pipeline(tmp_path, oracle):
    """README's DNA route: fagzToCompact4 (forward + reverse complement, N runs replaced) -> bwtb3m inputtype=compactstream
    -> checkbwt -i compactstream; the .bwt equals the naive BWT of the symbols the oracle reads from the container."""
    import gzip
    from bwtb3m_b200 import files
    rng = np.random.default_rng(23)
    lines = []
    for k, l in enumerate((30_001, 12_345, 77)):
        seq = "".join(rng.choice(list("ACGT"), size=l))
        if k == 0:
            seq = seq[:1000] + "N" * 500 + seq[1500:]
        lines.append(">seq%d" % k)
        lines.extend(seq[i:i + 70] for i in range(0, l, 70))
    fa = tmp_path / "g.fa.gz"
    fa.write_bytes(gzip.compress(("\n".join(lines) + "\n").encode()))
    r = subprocess.run([os.path.join(BIN, "fagzToCompact4"), "verbose=0", str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cfn = tmp_path / "g.compact"
    t, bits = oracle.decode_compact(cfn.read_bytes())
    assert bits == 2 and t.size == 2 * (30_001 + 12_345 + 77)
    sa = oracle.sa_circular(t)
    bwt, isa = oracle.bwt_from_sa(t, sa)
    out = tmp_path / "g.bwt"
    r = subprocess.run([os.path.join(BIN, "bwtb3m"), "inputtype=compactstream", "outputfilename=" + str(out), "sasamplingrate=16",
                        "isasamplingrate=64", str(cfn)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(files.read_bwt(str(out)), bwt)
    rate, v = files.read_sampled(str(tmp_path / "g.sa"))
    assert rate == 16 and np.array_equal(v, sa[::16].astype(np.uint64))
    rate, v = files.read_sampled(str(tmp_path / "g.isa"))
    assert rate == 64 and np.array_equal(v, isa[::64].astype(np.uint64))
    # the verifier reads the same container (needs the anchors: rebuild with bwtonly=1, which keeps .preisa)
    res = files.compute_bwt(str(cfn), inputtype="compactstream", outputfilename=str(tmp_path / "v.bwt"), bwtonly=True)
    ok, bad = files.check_bwt(res["bwtfn"], str(cfn), inputtype="compactstream")
    assert ok and bad == 0
    r = subprocess.run([os.path.join(BIN, "checkbwt"), "-i", "compactstream", res["bwtfn"], str(cfn)], capture_output=True, text=True)
    assert r.returncode == 0 and "[V] gok=1" in r.stderr


def test_lf_speed_instrument(tmp_path, oracle):
    """bwttestdecodespeed on the GPU: runs from the files bwtb3m wrote and reports a positive rate."""
    from bwtb3m_b200 import files
    data, t = make_pac(oracle, 300_001, 12)
    fn = tmp_path / "g.pac"
    data.tofile(fn)
    res = files.compute_bwt(str(fn), inputtype="pacterm", outputfilename=str(tmp_path / "g.bwt"), isasamplingrate=64)
    sps, sec = files.lf_speed(res["bwtfn"], 4096, 200)
    assert sps > 1e6 and sec > 0
    r = subprocess.run([os.path.join(BIN, "bwttestdecodespeed"), res["bwtfn"], "chains=2048", "steps=100"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[0] == "2048"
