"""ctypes binding of the CPU oracle (oracle/libb3m_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package bwtb3m_b200 never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libb3m_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.orc_pac_numsyms.restype = C.c_uint64
        L.orc_pac_numsyms.argtypes = [u8p, C.c_uint64]
        L.orc_decode_pac.restype = C.c_uint64
        L.orc_decode_pac.argtypes = [u8p, C.c_uint64, C.c_int, u8p]
        L.orc_encode_pac.restype = C.c_uint64
        L.orc_encode_pac.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_decode_compact.restype = C.c_uint64
        L.orc_decode_compact.argtypes = [u8p, C.c_uint64, u8p]
        L.orc_naive_rotation_sort.restype = None
        L.orc_naive_rotation_sort.argtypes = [u8p, C.c_uint64, u32p]
        L.orc_bwt_from_sa.restype = None
        L.orc_bwt_from_sa.argtypes = [u8p, C.c_uint64, u32p, u8p, u32p]
        L.orc_sa_circular.restype = C.c_int
        L.orc_sa_circular.argtypes = [u8p, C.c_uint64, u32p]
        L.orc_checkbwt.restype = C.c_int
        L.orc_checkbwt.argtypes = [u8p, C.c_uint64, u8p, u64p, C.c_uint64, C.c_uint, u64p]
        L.orc_default_numblocks.restype = C.c_uint64
        L.orc_default_numblocks.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_b3m.restype = C.c_int
        L.orc_b3m.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint, u8p, u64p, u64p]
        L.orc_ssa.restype = C.c_int
        L.orc_ssa.argtypes = [u8p, C.c_uint64, u64p, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, C.c_uint]
        L.orc_lf_speed.restype = C.c_double
        L.orc_lf_speed.argtypes = [u8p, C.c_uint64, u64p, C.c_uint64, C.c_uint, C.c_uint64, u64p]
        L.orc_to_bwa.restype = C.c_int
        L.orc_to_bwa.argtypes = [u8p, C.c_uint64, u64p, C.c_uint64, u8p, u64p, u8p, u64p]
        L.orc_max_threads.restype = C.c_uint
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def max_threads():
    return int(lib().orc_max_threads())


def decode_pac(filebytes, term):
    f = _u8(np.frombuffer(bytes(filebytes), dtype=np.uint8))
    l = lib().orc_pac_numsyms(_p(f, C.c_uint8), f.size)
    out = np.empty(l + (1 if term else 0), dtype=np.uint8)
    n = lib().orc_decode_pac(_p(f, C.c_uint8), f.size, 1 if term else 0, _p(out, C.c_uint8))
    assert n == out.size
    return out


def decode_compact(filebytes):
    """Symbols (one per byte) and bits per symbol of a compactstream container, either byte order
    [layout unpinned; the reader the reference uses is CompactDecoderWrapper, decodecompact.cpp:30-36]."""
    f = _u8(np.frombuffer(bytes(filebytes), dtype=np.uint8))
    assert f.size >= 32
    b = int.from_bytes(f[0:8].tobytes(), "big")
    n = int.from_bytes(f[8:16].tobytes(), "big")
    if not 1 <= b <= 64:
        b = int.from_bytes(f[0:8].tobytes(), "little")
        n = int.from_bytes(f[8:16].tobytes(), "little")
    out = np.empty(n, dtype=np.uint8)
    got = lib().orc_decode_compact(_p(f, C.c_uint8), f.size, _p(out, C.c_uint8))
    assert got == n
    return out, b


def encode_compact(syms, bits, layout="le"):
    """The container CompactArrayWriterFile(fn, bits) leaves behind (fagzToCompact4.cpp:105,232,265), as numpy
    bit arithmetic independent of the product's writer: 4 uint64 (bits, n, words, words), then 64-bit words with the
    symbols MSB first, zero padded.  layout "le": native little-endian numbers and words; "be": big-endian."""
    s = _u8(syms)
    assert 1 <= bits <= 8 and (s.size == 0 or int(s.max()) < (1 << bits)) and layout in ("le", "be")
    allbits = np.unpackbits(s.reshape(-1, 1), axis=1)[:, 8 - bits:].reshape(-1)
    words = (s.size * bits + 63) // 64
    pad = words * 64 - allbits.size
    payload = np.packbits(np.concatenate([allbits, np.zeros(pad, dtype=np.uint8)]))  # big-endian bit stream
    order = "big" if layout == "be" else "little"
    if layout == "le":
        payload = payload.reshape(-1, 8)[:, ::-1].reshape(-1)
    hdr = b"".join(int(v).to_bytes(8, order) for v in (bits, s.size, words, words))
    return np.frombuffer(hdr + payload.tobytes(), dtype=np.uint8).copy()


def encode_pac(syms):
    s = _u8(syms)
    out = np.zeros(s.size // 4 + 3, dtype=np.uint8)
    k = lib().orc_encode_pac(_p(s, C.c_uint8), s.size, _p(out, C.c_uint8))
    return out[:k].copy()


def naive_sa(text):
    t = _u8(text)
    sa = np.empty(t.size, dtype=np.uint32)
    lib().orc_naive_rotation_sort(_p(t, C.c_uint8), t.size, _p(sa, C.c_uint32))
    return sa


def sa_circular(text):
    t = _u8(text)
    sa = np.empty(t.size, dtype=np.uint32)
    rc = lib().orc_sa_circular(_p(t, C.c_uint8), t.size, _p(sa, C.c_uint32))
    assert rc == 0
    return sa


def bwt_from_sa(text, sa):
    t = _u8(text)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    bwt = np.empty(t.size, dtype=np.uint8)
    isa = np.empty(t.size, dtype=np.uint32)
    lib().orc_bwt_from_sa(_p(t, C.c_uint8), t.size, _p(sa, C.c_uint32), _p(bwt, C.c_uint8), _p(isa, C.c_uint32))
    return bwt, isa


def checkbwt(text, bwt, preisa_pairs, numthreads=8):
    """Restated checkbwt; preisa_pairs is a flat uint64 array (rank,pos,rank,pos,...)."""
    t, b = _u8(text), _u8(bwt)
    pp = np.ascontiguousarray(preisa_pairs, dtype=np.uint64).ravel()
    checked = C.c_uint64(0)
    rc = lib().orc_checkbwt(_p(t, C.c_uint8), t.size, _p(b, C.c_uint8), _p(pp, C.c_uint64), pp.size // 2,
                            numthreads, C.byref(checked))
    return rc, int(checked.value)


def default_numblocks(fs, mem, threads):
    return int(lib().orc_default_numblocks(fs, mem, threads))


def b3m(text, nblocks=1, rate=64, largelcpthres=16384, nthreads=1):
    """Block sort + gap + merge.  Returns (bwt, preisa_pairs[k,2] (rank,pos), stats dict)."""
    t = _u8(text)
    n = t.size
    bwt = np.empty(n, dtype=np.uint8)
    ns = (n + rate - 1) // rate
    pp = np.empty(2 * ns, dtype=np.uint64)
    st = np.zeros(4, dtype=np.uint64)
    rc = lib().orc_b3m(_p(t, C.c_uint8), n, nblocks, rate, largelcpthres, nthreads, _p(bwt, C.c_uint8),
                       _p(pp, C.c_uint64), _p(st, C.c_uint64))
    if rc != 0:
        raise RuntimeError("orc_b3m failed rc=%d" % rc)
    stats = {"gap_lf_steps": int(st[0]), "max_lcpnext": int(st[1]), "leaf_s": st[2] * 1e-6, "merge_s": st[3] * 1e-6}
    return bwt, pp.reshape(-1, 2), stats


def ssa(bwt, preisa_pairs, sarate=32, isarate=32, nthreads=1):
    b = _u8(bwt)
    n = b.size
    pp = np.ascontiguousarray(preisa_pairs, dtype=np.uint64).ravel()
    sa = np.empty((n + sarate - 1) // sarate, dtype=np.uint64)
    isa = np.empty((n + isarate - 1) // isarate, dtype=np.uint64)
    rc = lib().orc_ssa(_p(b, C.c_uint8), n, _p(pp, C.c_uint64), pp.size // 2, sarate, isarate,
                       _p(sa, C.c_uint64), _p(isa, C.c_uint64), nthreads)
    if rc != 0:
        raise RuntimeError("orc_ssa failed rc=%d" % rc)
    return sa, isa


def lf_speed(bwt, isa_samples, tpar=1, maxsteps=1 << 27):
    b = _u8(bwt)
    s = np.ascontiguousarray(isa_samples, dtype=np.uint64)
    cs = C.c_uint64(0)
    return float(lib().orc_lf_speed(_p(b, C.c_uint8), b.size, _p(s, C.c_uint64), s.size, tpar, maxsteps, C.byref(cs)))


def to_bwa(bwt, sa_samples, sarate):
    b = _u8(bwt)
    n = b.size
    s = np.ascontiguousarray(sa_samples, dtype=np.uint64)
    ob = np.zeros(40 + 4 * ((n - 1 + 15) // 16) + 8, dtype=np.uint8)
    osa = np.zeros(56 + 8 * s.size + 8, dtype=np.uint8)
    lb, ls = C.c_uint64(0), C.c_uint64(0)
    rc = lib().orc_to_bwa(_p(b, C.c_uint8), n, _p(s, C.c_uint64), sarate, _p(ob, C.c_uint8), C.byref(lb),
                          _p(osa, C.c_uint8), C.byref(ls))
    if rc != 0:
        raise RuntimeError("orc_to_bwa failed rc=%d" % rc)
    return ob[: lb.value].tobytes(), osa[: ls.value].tobytes()
