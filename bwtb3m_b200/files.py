"""Host-side mirror of the file-level C ABI (include/b3m.h): compute_bwt / compute_ssa / to_bwa
with the option names of the reference's command line (/root/reference/src/bwtb3m.cpp:43-56),
plus readers of the output files."""
import ctypes as C
import os

import numpy as np

from ._lib import Options, Result, lib
from .engine import B3MError


def _b(s):
    return None if s is None else os.fsencode(s)


def compute_bwt(fn, inputtype="bytestream", outputfilename=None, sasamplingrate=32, isasamplingrate=262144, mem=0,
                numthreads=0, bwtonly=False, tmpprefix=None, sparsetmpprefix=None, copyinputtomemory=False,
                largelcpthres=16384, verbose=0, device=0, numblocks=0, ngpus=1):
    """BwtMergeSort::computeBwt(options) -> dict of result file names (BwtMergeSortResult)."""
    L = lib()
    o = Options()
    L.b3m_options_init(C.byref(o))
    keep = [_b(fn), _b(inputtype), _b(outputfilename), _b(tmpprefix), _b(sparsetmpprefix)]
    o.fn, o.inputtype, o.outputfilename, o.tmpprefix, o.sparsetmpprefix = keep
    o.sasamplingrate, o.isasamplingrate, o.mem = sasamplingrate, isasamplingrate, mem
    if numthreads:
        o.numthreads = numthreads
    o.bwtonly = 1 if bwtonly else 0
    o.copyinputtomemory = 1 if copyinputtomemory else 0
    o.largelcpthres, o.verbose, o.device, o.numblocks = largelcpthres, verbose, device, numblocks
    o.ngpus = ngpus
    r = Result()
    err = C.create_string_buffer(2048)
    if L.b3m_compute_bwt(C.byref(o), C.byref(r), err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))
    out = {k: getattr(r, k).decode() for k in ("textfn", "bwtfn", "histfn", "preisafn", "metafn", "safn", "isafn")}
    out.update(n=int(r.n), numblocks=int(r.numblocks), seconds_total=float(r.seconds_total), seconds_device=float(r.seconds_device))
    return out


def compute_ssa(bwtfn, sasamplingrate=32, isasamplingrate=32, numthreads=0, verbose=0, ref_isa=None, ref_sa=None, device=0):
    """BwtComputeSSA::computeSSA (/root/reference/src/bwtcomputessa.cpp:39-51)."""
    err = C.create_string_buffer(2048)
    rc = lib().b3m_compute_ssa(_b(bwtfn), sasamplingrate, isasamplingrate, b"", 0, numthreads or (os.cpu_count() or 1), 2 << 30, 1024,
                               verbose, _b(ref_isa) or b"", _b(ref_sa) or b"", device, err, len(err))
    if rc != 0:
        raise B3MError(err.value.decode(errors="replace"))


def to_bwa(inbwt, outbwt, outsa):
    """MausFmToBwaConversion::rewrite (/root/reference/src/bwtb3mtobwa.cpp:29)."""
    err = C.create_string_buffer(2048)
    if lib().b3m_to_bwa(_b(inbwt), _b(outbwt), _b(outsa), err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))


def bwt_length(bwtfn):
    n = C.c_uint64(0)
    err = C.create_string_buffer(2048)
    if lib().b3m_bwt_length(_b(bwtfn), C.byref(n), err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))
    return int(n.value)


def read_bwt(bwtfn, numthreads=0):
    """Decoded symbols of a .bwt container (RLDecoder loop of bwtb3mdecoderl.cpp:27-46)."""
    n = bwt_length(bwtfn)
    out = np.empty(n, dtype=np.uint8)
    err = C.create_string_buffer(2048)
    if lib().b3m_bwt_decode(_b(bwtfn), C.c_void_p(out.ctypes.data), n, numthreads or (os.cpu_count() or 1), err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))
    return out


def write_bwt_host(bwtfn, syms):
    a = np.ascontiguousarray(syms, dtype=np.uint8)
    err = C.create_string_buffer(2048)
    if lib().b3m_bwt_encode_host(_b(bwtfn), C.c_void_p(a.ctypes.data), a.size, err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))


def block_sym_histograms(bwtfn, outfn, minsym, maxsym, numthreads=0):
    """Writes the `.sparserank` file of a .bwt (RLDecoder::getBlockSymHistograms, bwtdecodeblock.cpp:356-365); returns
    the number of blocks."""
    nb = C.c_uint64(0)
    err = C.create_string_buffer(2048)
    if lib().b3m_bwt_block_sym_histograms(_b(bwtfn), _b(outfn), minsym, maxsym, numthreads or (os.cpu_count() or 1), C.byref(nb), err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))
    return int(nb.value)


def bwt_rank(bwtfn, sparserankfn, minsym, maxsym, sym, i):
    """rank_sym(L, i) from the .bwt and its .sparserank file (SparseRank::rankm, bwtdecodeblock.cpp:210-242)."""
    r = C.c_uint64(0)
    err = C.create_string_buffer(2048)
    if lib().b3m_bwt_rank(_b(bwtfn), _b(sparserankfn), minsym, maxsym, sym, i, C.byref(r), err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))
    return int(r.value)


def write_compact(fn, syms, bits):
    """Compact container for inputtype=compactstream (CompactArrayWriterFile, fagzToCompact4.cpp:105,232,265)."""
    a = np.ascontiguousarray(syms, dtype=np.uint8)
    err = C.create_string_buffer(2048)
    if lib().b3m_compact_write(_b(fn), bits, C.c_void_p(a.ctypes.data), a.size, err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))


def read_compact(fn):
    """(symbols one per byte, bits per symbol) of a compact container (CompactDecoderWrapper, decodecompact.cpp:30-36)."""
    err = C.create_string_buffer(2048)
    bits, n = C.c_uint(0), C.c_uint64(0)
    if lib().b3m_compact_info(_b(fn), C.byref(bits), C.byref(n), err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))
    out = np.empty(n.value, dtype=np.uint8)
    if lib().b3m_compact_read(_b(fn), C.c_void_p(out.ctypes.data), n.value, err, len(err)) != 0:
        raise B3MError(err.value.decode(errors="replace"))
    return out, bits.value


def read_sampled(fn):
    """.sa / .isa: native uint64 [rate][count][values] (/root/reference/src/sasubsample.cpp:34-58)."""
    a = np.fromfile(fn, dtype=np.uint64)
    if a.size < 2 or a.size != 2 + int(a[1]):
        raise B3MError("malformed sampled array " + fn)
    return int(a[0]), a[2:]


def read_preisa(fn):
    """.preisa: native uint64 (rank,pos) pairs (/root/reference/src/hwtPreIsaToIsa.cpp:55-77)."""
    a = np.fromfile(fn, dtype=np.uint64)
    if a.size % 2:
        raise B3MError("malformed .preisa " + fn)
    return a.reshape(-1, 2)


def read_hist(fn):
    a = np.fromfile(fn, dtype=">u8")
    if a.size < 1 or a.size != 1 + 2 * int(a[0]):
        raise B3MError("malformed .hist " + fn)
    return {int(a[1 + 2 * i]): int(a[2 + 2 * i]) for i in range(int(a[0]))}


def check_bwt(bwtfn, textfn, inputtype="bytestream", numthreads=0, device=0, verbose=0):
    """checkbwt (reference: src/checkbwt.cpp:26-246) on the GPU; returns (ok, mismatches)."""
    ok, bad = C.c_int(0), C.c_uint64(0)
    err = C.create_string_buffer(2048)
    rc = lib().b3m_check_bwt(os.fsencode(bwtfn), os.fsencode(textfn), inputtype.encode(), numthreads or (os.cpu_count() or 1), device, verbose,
                             C.byref(ok), C.byref(bad), err, len(err))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return bool(ok.value), int(bad.value)


def lf_speed(bwtfn, nchains, steps=0, numthreads=0, device=0):
    """bwttestdecodespeed (reference: src/bwttestdecodespeed.cpp:27-97) on the GPU; returns (LF steps per second, seconds)."""
    sps, sec = C.c_double(0), C.c_double(0)
    err = C.create_string_buffer(2048)
    rc = lib().b3m_lf_speed(os.fsencode(bwtfn), nchains, steps, numthreads or (os.cpu_count() or 1), device, C.byref(sps), C.byref(sec), err, len(err))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return float(sps.value), float(sec.value)
