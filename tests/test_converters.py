"""Host-side input converters and the compactstream container: fagzToCompact4,
digitsToCompact, decodecompact, b3m_compact_* -- CPU only, checked against the oracle's reader and an
independent numpy writer.  Reference behaviour: /root/reference/src/fagzToCompact4.cpp:80-264,
/root/reference/src/digitsToCompact.cpp:26-92, /root/reference/src/decodecompact.cpp:21-44."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", ROOT, "all"])


@pytest.mark.parametrize("bits", [1, 2, 3, 4, 5, 7, 8])
@pytest.mark.parametrize("n", [0, 1, 21, 64, 1000, 70_001])
def test_compact_container_roundtrip_vs_oracle(tmp_path, oracle, bits, n):
    from bwtb3m_b200 import files
    t = np.random.default_rng(bits * 1000 + n).integers(0, 1 << bits, size=n, dtype=np.uint8)
    fn = str(tmp_path / "t.compact")
    files.write_compact(fn, t, bits)
    raw = np.fromfile(fn, dtype=np.uint8)
    # same bytes as the independent numpy writer (native little-endian words), same symbols through the oracle's reader and ours
    assert raw.tobytes() == oracle.encode_compact(t, bits, "le").tobytes()
    got, b = oracle.decode_compact(raw.tobytes())
    assert b == bits and np.array_equal(got, t)
    got2, b2 = files.read_compact(fn)
    assert b2 == bits and np.array_equal(got2, t)
    # the big-endian reading of the container is accepted by both readers too
    be = oracle.encode_compact(t, bits, "be")
    got, b = oracle.decode_compact(be.tobytes())
    assert b == bits and np.array_equal(got, t)
    be.tofile(fn)
    got2, b2 = files.read_compact(fn)
    assert b2 == bits and np.array_equal(got2, t)
    r = subprocess.run([os.path.join(BIN, "decodecompact"), fn], capture_output=True)
    assert r.returncode == 0 and r.stdout == t.tobytes()


def test_compact_container_errors(tmp_path):
    from bwtb3m_b200 import files
    from bwtb3m_b200.files import B3MError
    fn = str(tmp_path / "t.compact")
    with pytest.raises(B3MError):
        files.write_compact(fn, np.array([0, 4, 1], dtype=np.uint8), 2)  # symbol needs 3 bits
    with pytest.raises(B3MError):
        files.write_compact(fn, np.zeros(3, dtype=np.uint8), 9)
    files.write_compact(fn, np.arange(200, dtype=np.uint8) % 4, 2)
    raw = np.fromfile(fn, dtype=np.uint8)
    raw[:40].tofile(fn)  # truncated payload
    with pytest.raises(B3MError):
        files.read_compact(fn)
    raw[:8].tofile(fn)
    with pytest.raises(B3MError):
        files.read_compact(fn)


def _fasta(records, width=60):
    out = []
    for name, seq in records:
        out.append(">" + name)
        out.extend(seq[i:i + width] for i in range(0, len(seq), width))
    return ("\n".join(out) + "\n").encode()


def _meta(fn):
    a = np.fromfile(fn, dtype=">u8").astype(np.uint64).tolist()
    nseq, pos, recs = a[0], 1, []
    for _ in range(nseq):
        l, nr = a[pos], a[pos + 1]
        iv = [(a[pos + 2 + 2 * k], a[pos + 3 + 2 * k]) for k in range(nr)]
        pos += 2 + 2 * nr
        recs.append((l, iv))
    assert pos == len(a)
    return recs


@pytest.mark.parametrize("gz", [1, 0])
@pytest.mark.parametrize("rc", [1, 0])
def test_fagzToCompact4(tmp_path, oracle, gz, rc):
    rng = np.random.default_rng(17 + gz + 2 * rc)
    recs = []
    for name, l in (("chr1 first", 1000), ("chr2", 61), ("chr3", 1), ("chrN", 7)):
        recs.append((name, "".join(rng.choice(list("ACGTacgt"), size=l))))
    s0 = list(recs[0][1])
    s0[0:5] = "NNNNN"       # run at the start
    s0[300:420] = "n" * 120  # lower case too
    s0[999] = "R"            # any non-ACGT letter, at the end
    recs[0] = (recs[0][0], "".join(s0))
    recs[3] = (recs[3][0], "NNNNNNN")
    data = _fasta(recs)
    fa = tmp_path / ("in.fa.gz" if gz else "in.fa")
    fa.write_bytes(gzip.compress(data) if gz else data)
    r = subprocess.run([os.path.join(BIN, "fagzToCompact4"), "rc=%d" % rc, "gz=%d" % gz, str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = tmp_path / "in.compact"  # default name: input minus .gz/.fa plus .compact
    assert out.exists() and (tmp_path / "in.compact.meta").exists()
    assert "1 in chr1 first...done, input size" in r.stderr and "Done, total input size %d" % sum(len(s) + 1 for _, s in recs) in r.stderr
    syms, bits = oracle.decode_compact(out.read_bytes())
    assert bits == 2 and syms.size == sum(len(s) for _, s in recs) * (2 if rc else 1)
    meta = _meta(tmp_path / "in.compact.meta")
    assert [m[0] for m in meta] == [len(s) for _, s in recs]
    assert meta[0][1] == [(0, 5), (300, 420), (999, 1000)] and meta[1][1] == [] and meta[3][1] == [(0, 7)]
    code = {c: i for i, c in enumerate("ACGT")}
    pos = 0
    for (name, seq), (l, iv) in zip(recs, meta):
        fwd = syms[pos:pos + l]
        repl = np.zeros(l, dtype=bool)
        for a, b in iv:
            repl[a:b] = True
        want = np.array([code.get(c.upper(), 0) for c in seq], dtype=np.uint8)
        assert np.array_equal(fwd[~repl], want[~repl]) and fwd.max(initial=0) < 4
        pos += l
        if rc:
            assert np.array_equal(syms[pos:pos + l], 3 - fwd[::-1])  # reverse complement, replaced bases included
            pos += l
    assert pos == syms.size
    # the same file again: identical bytes (fixed-seed replacement bases); decodecompact agrees with the oracle
    r2 = subprocess.run([os.path.join(BIN, "fagzToCompact4"), "rc=%d" % rc, "gz=%d" % gz, "verbose=0",
                         "outputfilename=" + str(tmp_path / "again.compact"), str(fa)], capture_output=True, text=True)
    assert r2.returncode == 0 and "chr1" not in r2.stderr
    assert (tmp_path / "again.compact").read_bytes() == out.read_bytes()
    d = subprocess.run([os.path.join(BIN, "decodecompact"), str(out), str(out)], capture_output=True)
    assert d.returncode == 0 and d.stdout == syms.tobytes() * 2


def test_fagzToCompact4_list_file_and_errors(tmp_path, oracle):
    a, b = tmp_path / "x_a.fa", tmp_path / "x_b.fa"
    a.write_bytes(_fasta([("s1", "ACGT" * 10)]))
    b.write_bytes(_fasta([("s2", "TTGCA")]))
    lst = tmp_path / "names.txt"
    lst.write_text("%s\n\n%s\n" % (a, b))
    r = subprocess.run([os.path.join(BIN, "fagzToCompact4"), "gz=0", "rc=0", "inputfilenames=" + str(lst)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = tmp_path / "x_.compact"  # longest common prefix of the input names
    syms, _ = oracle.decode_compact(out.read_bytes())
    assert syms.tolist() == [0, 1, 2, 3] * 10 + [3, 3, 2, 1, 0]
    r = subprocess.run([os.path.join(BIN, "fagzToCompact4"), "gz=0", str(tmp_path / "missing.fa")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr
    r = subprocess.run([os.path.join(BIN, "fagzToCompact4")], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    r = subprocess.run([os.path.join(BIN, "decodecompact"), str(a)], capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.parametrize("term", [0, 1])
@pytest.mark.parametrize("gz", [0, 1])
def test_digitsToCompact(tmp_path, oracle, term, gz):
    digits = np.random.default_rng(5).integers(0, 10, size=20_001, dtype=np.uint8)
    text = (digits + ord("0")).astype(np.uint8).tobytes()
    out = tmp_path / "d.compact"
    r = subprocess.run([os.path.join(BIN, "digitsToCompact"), "term=%d" % term, "gz=%d" % gz, "outputfilename=" + str(out)],
                       input=gzip.compress(text) if gz else text, capture_output=True)
    assert r.returncode == 0, r.stderr
    syms, bits = oracle.decode_compact(out.read_bytes())
    assert bits == 4
    want = np.concatenate([digits + 1, [0]]).astype(np.uint8) if term else digits
    assert np.array_equal(syms, want)
    r = subprocess.run([os.path.join(BIN, "digitsToCompact"), "outputfilename=" + str(out)], input=b"123\n", capture_output=True, text=False)
    assert r.returncode == 1 and b"non decimal digit" in r.stderr


def _revcomp_codes(codes, a, t, other):
    """reverse complement on mapped symbols: A..T = a..t swap ends, `other` stays"""
    c = codes[::-1].copy()
    m = (c >= a) & (c <= t)
    c[m] = (a + t) - c[m]
    return c


@pytest.mark.parametrize("rc", [1, 0])
def test_fagzToCompact_3bit(tmp_path, oracle, rc):
    """3 bit per symbol: A,C,G,T = 1..4, other letters 5, a 0 behind every sequence (fagzToCompact.cpp:108-160)."""
    recs = [("s1 x", "ACGTNacgtRY"), ("s2", "T"), ("s3", "GGGCCCAAATTT" * 9)]
    fa = tmp_path / "in.fa.gz"
    fa.write_bytes(gzip.compress(_fasta(recs, width=13)))
    out = tmp_path / "o.compact"
    r = subprocess.run([os.path.join(BIN, "fagzToCompact"), "rc=%d" % rc, "outputfilename=" + str(out), str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Done, total input size %d" % sum(len(s) + 1 for _, s in recs) in r.stderr
    syms, bits = oracle.decode_compact(out.read_bytes())
    assert bits == 3
    code = {"A": 1, "C": 2, "G": 3, "T": 4}
    want = []
    for _, s in recs:
        f = np.array([code.get(c.upper(), 5) for c in s], dtype=np.uint8)
        want += f.tolist() + [0]
        if rc:
            want += _revcomp_codes(f, 1, 4, 5).tolist() + [0]
    assert syms.tolist() == want
    # limit=: files are only opened while the accumulated size is below it
    r = subprocess.run([os.path.join(BIN, "fagzToCompact"), "rc=0", "verbose=0", "limit=5", "outputfilename=" + str(out), str(fa), str(fa)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms2, _ = oracle.decode_compact(out.read_bytes())
    assert syms2.size == sum(len(s) + 1 for _, s in recs)  # the second file is skipped


@pytest.mark.parametrize("rc", [1, 0])
def test_fagzToCompactUTerm(tmp_path, oracle, rc):
    """A,C,G,T = 2..5, other letters 6, and behind sequence k its id in `seqbits` symbols 0/1, MSB first
    (fagzToCompactUTerm.cpp:78-85,140-210)."""
    recs = [("a", "ACGTN"), ("b", "tt"), ("c", "GATTACA")]
    fa = tmp_path / "in.fa"
    fa.write_bytes(_fasta(recs))
    out = tmp_path / "u.compact"
    r = subprocess.run([os.path.join(BIN, "fagzToCompactUTerm"), "gz=0", "rc=%d" % rc, "outputfilename=" + str(out), str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    numseq = len(recs) * (2 if rc else 1)
    assert "[V] numseq=%d" % numseq in r.stderr
    seqbits = (numseq - 1).bit_length()
    syms, bits = oracle.decode_compact(out.read_bytes())
    assert bits == 3
    code = {"A": 2, "C": 3, "G": 4, "T": 5}
    want, sid = [], 0
    for _, s in recs:
        f = np.array([code.get(c.upper(), 6) for c in s], dtype=np.uint8)
        for part in ([f, _revcomp_codes(f, 2, 5, 6)] if rc else [f]):
            want += part.tolist() + [(sid >> (seqbits - 1 - i)) & 1 for i in range(seqbits)]
            sid += 1
    assert syms.tolist() == want
