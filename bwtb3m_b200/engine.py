"""Host-side mirror of the engine-level C ABI (include/b3m.h): one Engine per GPU."""
import ctypes as C

import numpy as np

from ._lib import INPUT_TYPES, BuildParams, Info, lib


SORTPATHS = {"auto": 0, "lsd": 1, "msd": 2}  # B3M_SORT_* of include/b3m.h


class B3MError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Engine:
    """Drives K1..K7 on one GPU through libb3m.so.  Mirrors the argument meaning of the
    reference's BwtMergeSortOptions (/root/reference/src/bwtb3m.cpp:43-56)."""

    def __init__(self, device=0, stream=None):
        self._lib = lib()
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = self._lib.b3m_engine_create(device, C.c_void_p(stream) if stream else None, C.byref(h), err, 1024)
        if rc != 0:
            raise B3MError(err.value.decode() or "b3m_engine_create failed (%d)" % rc)
        self._h = h
        self.device = device
        self.stream_ptr = stream or 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b3m_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise B3MError(self._lib.b3m_engine_last_error(self._h).decode())

    def load_host(self, data, inputtype="bytestream"):
        """data: bytes-like or uint8 numpy array holding the input FILE contents."""
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        self._keep = a
        self._check(self._lib.b3m_engine_load_host(self._h, C.c_void_p(a.ctypes.data), a.size, INPUT_TYPES[inputtype]))

    def load_host_ptr(self, ptr, nbytes, inputtype="bytestream"):
        self._check(self._lib.b3m_engine_load_host(self._h, C.c_void_p(ptr), nbytes, INPUT_TYPES[inputtype]))

    def load_device(self, dptr, nbytes, inputtype="bytestream"):
        self._check(self._lib.b3m_engine_load_device(self._h, C.c_void_p(dptr), nbytes, INPUT_TYPES[inputtype]))

    def build(self, numblocks=1, preisarate=0, sasamplingrate=32, isasamplingrate=262144, bwtonly=False,
              largelcpthres=16384, sampling="auto", host_sa_ptr=0, host_bwa_ptr=0, sortpath="auto", gapmode="auto"):
        """sampling: "auto" (sampled SA/ISA straight from the suffix array when one block holds the whole
        text, LF walk otherwise) or "walk" (always the reference's LF walk from the anchors).
        gapmode: how K5 counts a gap array: "auto" (by its size), "atomic" or "list" (include/b3m.h B3M_GAP_*)."""
        p = BuildParams(numblocks, preisarate, sasamplingrate, isasamplingrate, 1 if bwtonly else 0, largelcpthres,
                        {"auto": 0, "walk": 1}[sampling], host_sa_ptr or None, host_bwa_ptr or None, SORTPATHS[sortpath],
                        {"auto": 0, "atomic": 1, "list": 2}[gapmode])
        self._check(self._lib.b3m_engine_build(self._h, C.byref(p)))

    def info(self):
        i = Info()
        self._check(self._lib.b3m_engine_info(self._h, C.byref(i)))
        d = {k: getattr(i, k) for k, _ in Info._fields_ if k != "hist"}
        d["hist"] = {s: int(c) for s, c in enumerate(i.hist) if c}
        return d

    def fetch(self, bwt=True, preisa=True, sa=True, isa=True, out=None):
        """Returns dict with numpy arrays; `out` may hold preallocated (e.g. pinned) arrays."""
        i = self.info()
        out = dict(out or {})
        res = {}
        if bwt:
            res["bwt"] = out.get("bwt") if out.get("bwt") is not None else np.empty(i["n"], dtype=np.uint8)
        if preisa:
            res["preisa"] = out.get("preisa") if out.get("preisa") is not None else np.empty((i["npreisa"], 2), dtype=np.uint64)
        if sa and i["nsa"]:
            res["sa"] = out.get("sa") if out.get("sa") is not None else np.empty(i["nsa"], dtype=np.uint64)
        if isa and i["nisa"]:
            res["isa"] = out.get("isa") if out.get("isa") is not None else np.empty(i["nisa"], dtype=np.uint64)
        self._check(self._lib.b3m_engine_fetch(self._h, _ptr(res.get("bwt")), _ptr(res.get("preisa")), _ptr(res.get("sa")),
                                               _ptr(res.get("isa"))))
        return res

    def fetch_ptrs(self, bwt_ptr, preisa_ptr, sa_ptr, isa_ptr):
        self._check(self._lib.b3m_engine_fetch(self._h, C.c_void_p(bwt_ptr) if bwt_ptr else None,
                                               C.c_void_p(preisa_ptr) if preisa_ptr else None,
                                               C.c_void_p(sa_ptr) if sa_ptr else None,
                                               C.c_void_p(isa_ptr) if isa_ptr else None))

    def write_bwt(self, fn):
        """K8: device-side run-length Huffman encoding of the BWT of the last build into `fn`."""
        import os
        self._check(self._lib.b3m_engine_write_bwt(self._h, os.fsencode(fn)))

    def fetch_runs(self):
        """(symbols uint8, lengths uint64) of the runs of the BWT."""
        n = C.c_uint64(0)
        self._check(self._lib.b3m_engine_fetch_runs(self._h, None, None, 0, C.byref(n)))
        syms = np.empty(n.value, dtype=np.uint8)
        lens = np.empty(n.value, dtype=np.uint64)
        self._check(self._lib.b3m_engine_fetch_runs(self._h, _ptr(syms), _ptr(lens), n.value, C.byref(n)))
        return syms, lens

    def shard_build(self, part, nparts, bwt_ptr, prerank_ptr, sa_ptr, isa_ptr, special_ptr, preisarate=0, sasamplingrate=32,
                    isasamplingrate=262144, bwtonly=False, sortpath="auto"):
        """Suffix-range sharding: sorts key range `part` of `nparts` into caller-owned, zeroed device
        buffers (global places); returns the number of suffixes this path left unresolved."""
        p = BuildParams(nparts, preisarate, sasamplingrate, isasamplingrate, 1 if bwtonly else 0, 16384, 0, None, None, SORTPATHS[sortpath])
        un = C.c_uint64(0)
        vp = lambda a: C.c_void_p(a) if a else None
        self._check(self._lib.b3m_engine_shard_build(self._h, part, nparts, C.byref(p), vp(bwt_ptr), vp(prerank_ptr), vp(sa_ptr), vp(isa_ptr),
                                                     vp(special_ptr), C.byref(un)))
        return int(un.value)

    def shard_rows(self, nparts):
        """first BWT row of every key range (nparts + 1 values)"""
        a = (C.c_uint64 * (nparts + 1))()
        self._check(self._lib.b3m_engine_shard_rows(self._h, nparts, a))
        return [int(x) for x in a]

    def pack_rows(self, rows_ptr, nrows, packed_ptr):
        """2-bit transport packing of BWT rows; returns False when the alphabet has more than four codes."""
        return self._lib.b3m_engine_pack_rows(self._h, C.c_void_p(rows_ptr), nrows, C.c_void_p(packed_ptr)) == 0

    def unpack_rows(self, packed_ptr, nrows, rows_ptr):
        self._check(self._lib.b3m_engine_unpack_rows(self._h, C.c_void_p(packed_ptr), nrows, C.c_void_p(rows_ptr)))

    def shard_finish(self, nparts, bwt_ptr, prerank_ptr, sa_ptr, isa_ptr, special_ptr):
        vp = lambda a: C.c_void_p(a) if a else None
        self._check(self._lib.b3m_engine_shard_finish(self._h, vp(bwt_ptr), vp(prerank_ptr), vp(sa_ptr), vp(isa_ptr), vp(special_ptr), nparts))

    def xshard_count(self, part, nparts, totals_ptr, preisarate=0, sasamplingrate=32, isasamplingrate=262144, bwtonly=False, sortpath="auto"):
        """Position-sharded build, step 1: counts of this part's text positions per leading-bits bin into the device buffer
        totals_ptr (2048 uint64 at most); returns the number of bins, 0 when the sorter does not apply to the text."""
        p = BuildParams(nparts, preisarate, sasamplingrate, isasamplingrate, 1 if bwtonly else 0, 16384, 0, None, None, SORTPATHS[sortpath])
        nb = C.c_uint32(0)
        self._check(self._lib.b3m_engine_xshard_count(self._h, part, nparts, C.byref(p), C.c_void_p(totals_ptr), C.byref(nb)))
        return int(nb.value)

    def xshard_scatter(self, all_totals, recs_ptrs, caps):
        """Step 2: all_totals = host uint64 array [nparts][nbins]; recs_ptrs[q] = part q's record array as mapped in this
        process; caps[q] = its capacity in records."""
        t = np.ascontiguousarray(all_totals, dtype=np.uint64)
        n = len(recs_ptrs)
        pa = (C.c_void_p * n)(*[C.c_void_p(x) for x in recs_ptrs])
        ca = (C.c_uint64 * n)(*caps)
        self._check(self._lib.b3m_engine_xshard_scatter(self._h, C.c_void_p(t.ctypes.data), pa, ca))

    def xshard_stream_sa(self, d_sa_local_ptr, host_sa_ptr):
        """Arms the next xshard_finish: this rank's SA samples also go to `d_sa_local_ptr` (this GPU) and from there
        into the page-locked host buffer `host_sa_ptr` while the finish kernel runs (b3m_engine_xshard_stream_sa)."""
        self._check(self._lib.b3m_engine_xshard_stream_sa(self._h, C.c_void_p(d_sa_local_ptr), C.c_void_p(host_sa_ptr)))

    def xshard_sa_delivered(self):
        d = C.c_int(0)
        self._check(self._lib.b3m_engine_xshard_sa_delivered(self._h, C.byref(d)))
        return bool(d.value)

    def xshard_finish(self, recs_own_ptr, bwt_ptr, prerank_ptr, sa_ptr, isa_ptr, special_ptr):
        """Step 3 (after every part has scattered): sorts this part's key range, outputs at their global places; returns the
        number of suffixes left unresolved."""
        un = C.c_uint64(0)
        vp = lambda a: C.c_void_p(a) if a else None
        self._check(self._lib.b3m_engine_xshard_finish(self._h, vp(recs_own_ptr), vp(bwt_ptr), vp(prerank_ptr), vp(sa_ptr), vp(isa_ptr), vp(special_ptr),
                                                       C.byref(un)))
        return int(un.value)

    def shard_adopt(self, nparts, bwt_ptr, prerank_ptr, sa_ptr, isa_ptr, special_ptr):
        """shard_finish without the copies: the engine refers to the caller's buffers."""
        vp = lambda a: C.c_void_p(a) if a else None
        self._check(self._lib.b3m_engine_shard_adopt(self._h, vp(bwt_ptr), vp(prerank_ptr), vp(sa_ptr), vp(isa_ptr), vp(special_ptr), nparts))

    def fetch_bwa(self, out=None, out_ptr=0):
        """BWA's packed BWT of the last pacterm build: (words uint32, primary, L2[5], seq_len).
        `out` (numpy uint32) or `out_ptr` (address of a buffer of enough words) may receive the words."""
        primary, seq_len = C.c_uint64(0), C.c_uint64(0)
        l2 = (C.c_uint64 * 5)()
        self._check(self._lib.b3m_engine_fetch_bwa(self._h, None, 0, C.byref(primary), l2, C.byref(seq_len)))
        nw = (seq_len.value + 15) >> 4
        if out_ptr:
            self._check(self._lib.b3m_engine_fetch_bwa(self._h, C.c_void_p(out_ptr), nw, None, None, None))
            words = None
        else:
            words = out if out is not None else np.empty(nw, dtype=np.uint32)
            self._check(self._lib.b3m_engine_fetch_bwa(self._h, _ptr(words), words.size, None, None, None))
        return words, int(primary.value), [int(x) for x in l2], int(seq_len.value)

    def pack_bwa(self, d_words_ptr, w_lo, w_hi):
        """K9 into a caller-owned device buffer: BWA's words [w_lo, w_hi) of the last pacterm build (no host copy)."""
        self._check(self._lib.b3m_engine_pack_bwa(self._h, C.c_void_p(d_words_ptr), w_lo, w_hi))

    def ssa_from_bwt(self, bwt, preisa_pairs, sasamplingrate=32, isasamplingrate=32):
        """K4 + K7 on an existing BWT (bwtcomputessa path); returns (sa, isa) samples."""
        b = np.ascontiguousarray(bwt, dtype=np.uint8)
        pp = np.ascontiguousarray(preisa_pairs, dtype=np.uint64).ravel()
        self._check(self._lib.b3m_engine_ssa_from_bwt(self._h, _ptr(b), b.size, _ptr(pp), pp.size // 2, sasamplingrate, isasamplingrate))
        i = self.info()
        sa = np.empty(i["nsa"], dtype=np.uint64)
        isa = np.empty(i["nisa"], dtype=np.uint64)
        self._check(self._lib.b3m_engine_fetch(self._h, None, None, _ptr(sa), _ptr(isa)))
        return sa, isa

    def default_preisarate(self, bwtonly=False):
        """Anchor spacing the engine would choose for the loaded text."""
        r = C.c_uint64(0)
        self._check(self._lib.b3m_engine_default_preisarate(self._h, 1 if bwtonly else 0, C.byref(r)))
        return int(r.value)

    def lf_bench(self, nchains, steps):
        ms = C.c_float(0)
        cs = C.c_uint64(0)
        self._check(self._lib.b3m_engine_lf_bench(self._h, nchains, steps, C.byref(ms), C.byref(cs)))
        return float(ms.value), int(cs.value)

    def set_profile(self, on=True):
        self._check(self._lib.b3m_engine_set_profile(self._h, 1 if on else 0))

    def kernel_times(self):
        """{name: {"launches": L, "ms": total ms, "bytes": algorithmic bytes}} since the last call."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self._lib.b3m_engine_kernel_times(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms, nbytes = line.rsplit(" ", 3)
            out[name] = {"launches": int(cnt), "ms": float(ms), "bytes": int(nbytes)}
        return out

    def sync(self):
        self._check(self._lib.b3m_engine_sync(self._h))


class DeviceMemory:
    """cudaMalloc'ed device memory and its CUDA IPC handle (b3m_dev_alloc / b3m_ipc_*): result buffers that the
    other processes of a multi-GPU build write directly."""

    def __init__(self, device):
        self._lib = lib()
        self.device = device

    def _call(self, fn, *args):
        err = C.create_string_buffer(512)
        if fn(*args, err, 512) != 0:
            raise B3MError(err.value.decode())

    def alloc(self, nbytes):
        p = C.c_void_p()
        self._call(self._lib.b3m_dev_alloc, self.device, nbytes, C.byref(p))
        return p.value

    def free(self, ptr):
        self._call(self._lib.b3m_dev_free, self.device, C.c_void_p(ptr))

    def export(self, ptr):
        h = C.create_string_buffer(64)
        self._call(self._lib.b3m_ipc_export, self.device, C.c_void_p(ptr), h)
        return h.raw

    def open(self, handle):
        p = C.c_void_p()
        self._call(self._lib.b3m_ipc_open, self.device, C.c_char_p(handle), C.byref(p))
        return p.value

    def close(self, ptr):
        self._call(self._lib.b3m_ipc_close, self.device, C.c_void_p(ptr))

    def copy(self, dst, src, nbytes, stream_ptr=0):
        """cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault) on `stream_ptr` of this device."""
        self._call(self._lib.b3m_dev_copy, self.device, C.c_void_p(dst), C.c_void_p(src), nbytes, C.c_void_p(stream_ptr) if stream_ptr else None)


class MultiEngine:
    """The multi-GPU build inside one process (C ABI b3m_multi_*, what `ngpus=` of bwtb3m selects): `ngpus`
    engines, one host thread per GPU inside libb3m.so.  After build() the results are read through
    `self.engine` (GPU 0's engine): fetch / fetch_bwa / write_bwt / info as after a single-GPU build."""

    def __init__(self, ngpus, devices=None):
        self._lib = lib()
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        devs = (C.c_int * ngpus)(*devices) if devices is not None else None
        rc = self._lib.b3m_multi_create(ngpus, devs, C.byref(h), err, 1024)
        if rc != 0:
            raise B3MError(err.value.decode() or "b3m_multi_create failed (%d)" % rc)
        self._h = h
        self.ngpus = ngpus
        # a non-owning view of engine 0
        self.engine = Engine.__new__(Engine)
        self.engine._lib = self._lib
        self.engine._h = None
        self.engine._view = C.c_void_p(self._lib.b3m_multi_engine(self._h, 0))
        self.engine.device = devices[0] if devices is not None else 0
        self.engine.stream_ptr = 0
        self.engine._h = self.engine._view
        self.engine.close = lambda: None

    def close(self):
        if getattr(self, "_h", None):
            self.engine._h = None
            self._lib.b3m_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise B3MError(self._lib.b3m_multi_last_error(self._h).decode())

    def load_host(self, data, inputtype="bytestream"):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        self._keep = a
        self._check(self._lib.b3m_multi_load_host(self._h, C.c_void_p(a.ctypes.data), a.size, INPUT_TYPES[inputtype]))

    def load_host_ptr(self, ptr, nbytes, inputtype="bytestream"):
        self._check(self._lib.b3m_multi_load_host(self._h, C.c_void_p(ptr), nbytes, INPUT_TYPES[inputtype]))

    def build(self, numblocks=1, preisarate=0, sasamplingrate=32, isasamplingrate=262144, bwtonly=False,
              largelcpthres=16384, sampling="auto", sortpath="auto"):
        p = BuildParams(numblocks, preisarate, sasamplingrate, isasamplingrate, 1 if bwtonly else 0, largelcpthres,
                        {"auto": 0, "walk": 1}[sampling], None, None, SORTPATHS[sortpath])
        self._check(self._lib.b3m_multi_build(self._h, C.byref(p)))

    def stats(self):
        s = C.create_string_buffer(128)
        a, b = C.c_double(0), C.c_double(0)
        self._check(self._lib.b3m_multi_stats(self._h, s, 128, C.byref(a), C.byref(b)))
        return {"strategy": s.value.decode(), "ms_load": a.value, "ms_build": b.value}
