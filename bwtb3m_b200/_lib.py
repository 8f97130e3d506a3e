"""Loads libb3m.so (the C ABI of include/b3m.h) with ctypes.  There is no fallback: if the
CUDA library is missing the import fails."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb3m.so")

INPUT_TYPES = {"bytestream": 0, "compactstream": 1, "pac": 2, "pacterm": 3}


class BuildParams(C.Structure):
    _fields_ = [
        ("numblocks", C.c_uint64),
        ("preisarate", C.c_uint64),
        ("sasamplingrate", C.c_uint64),
        ("isasamplingrate", C.c_uint64),
        ("bwtonly", C.c_int),
        ("largelcpthres", C.c_uint64),
        ("sampling", C.c_int),
        ("host_sa", C.c_void_p),
        ("host_bwa", C.c_void_p),
        ("sortpath", C.c_int),
        ("gapmode", C.c_int),
    ]


class Info(C.Structure):
    _fields_ = (
        [(k, C.c_uint64) for k in ("n", "sigma", "numblocks", "preisarate", "npreisa", "sasamplingrate", "nsa",
                                   "isasamplingrate", "nisa")]
        + [("hist", C.c_uint64 * 256)]
        + [(k, C.c_uint64) for k in ("sort_rounds", "radix_passes", "radix_bytes", "sort_active_sum", "sort_other_bytes",
                                     "gap_lf_steps", "walk_lf_steps", "walk_chains", "gap_chains", "merge_bytes",
                                     "extract_bytes", "dict_bytes", "decode_bytes", "launches", "max_lcpnext")]
        + [(k, C.c_float) for k in ("ms_decode", "ms_sort", "ms_extract", "ms_dict", "ms_gap", "ms_merge", "ms_walk",
                                    "ms_total")]
        + [(k, C.c_uint64) for k in ("sort_tied0", "sort_unresolved0", "arena_capacity", "arena_peak")]
    )


class Options(C.Structure):
    _fields_ = [
        ("fn", C.c_char_p), ("inputtype", C.c_char_p), ("outputfilename", C.c_char_p),
        ("sasamplingrate", C.c_uint64), ("isasamplingrate", C.c_uint64), ("mem", C.c_uint64),
        ("numthreads", C.c_uint64), ("bwtonly", C.c_int), ("tmpprefix", C.c_char_p),
        ("sparsetmpprefix", C.c_char_p), ("copyinputtomemory", C.c_int), ("largelcpthres", C.c_uint64),
        ("verbose", C.c_int), ("device", C.c_int), ("numblocks", C.c_uint64), ("ngpus", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [(k, C.c_char * 1024) for k in ("textfn", "bwtfn", "histfn", "preisafn", "metafn", "safn", "isafn")] + [
        ("n", C.c_uint64), ("numblocks", C.c_uint64), ("seconds_total", C.c_double), ("seconds_device", C.c_double)]


# every symbol include/b3m.h declares
EXPORTS = [
    "b3m_version", "b3m_parse_inputtype", "b3m_options_init", "b3m_compute_bwt", "b3m_compute_ssa", "b3m_to_bwa", "b3m_check_bwt", "b3m_lf_speed",
    "b3m_engine_create", "b3m_engine_destroy", "b3m_engine_last_error", "b3m_engine_load_host",
    "b3m_engine_load_device", "b3m_engine_build", "b3m_engine_info", "b3m_engine_fetch",
    "b3m_engine_device_results", "b3m_engine_lf_bench", "b3m_engine_sync", "b3m_engine_set_profile",
    "b3m_engine_kernel_times", "b3m_engine_write_bwt", "b3m_engine_fetch_runs", "b3m_engine_ssa_from_bwt",
    "b3m_bwt_length", "b3m_bwt_decode", "b3m_bwt_encode_host", "b3m_bwt_block_sym_histograms", "b3m_bwt_rank",
    "b3m_compact_write", "b3m_compact_info", "b3m_compact_read",
    "b3m_engine_blk_begin", "b3m_engine_blk_build_range", "b3m_engine_blk_chains", "b3m_engine_blk_zranks", "b3m_engine_blk_gap",
    "b3m_engine_blk_merge", "b3m_engine_blk_merge_samples", "b3m_engine_blk_finish",
    "b3m_engine_default_preisarate", "b3m_engine_fetch_bwa",
    "b3m_engine_shard_build", "b3m_engine_shard_finish", "b3m_engine_shard_rows", "b3m_engine_pack_rows", "b3m_engine_unpack_rows",
    "b3m_engine_shard_adopt", "b3m_engine_xshard_count", "b3m_engine_xshard_scatter", "b3m_engine_xshard_finish", "b3m_engine_xshard_stream_sa", "b3m_engine_xshard_sa_delivered", "b3m_dev_alloc", "b3m_dev_free", "b3m_ipc_export", "b3m_ipc_open", "b3m_ipc_close", "b3m_dev_copy", "b3m_engine_pack_bwa",
    "b3m_multi_create", "b3m_multi_destroy", "b3m_multi_last_error", "b3m_multi_load_host", "b3m_multi_build", "b3m_multi_engine", "b3m_multi_stats",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make` (nvcc, sm_100a); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, u8p, u64p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    L.b3m_version.restype = C.c_char_p
    L.b3m_parse_inputtype.argtypes = [C.c_char_p]
    L.b3m_engine_create.argtypes = [C.c_int, vp, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.b3m_engine_destroy.argtypes = [vp]
    L.b3m_engine_destroy.restype = None
    L.b3m_engine_last_error.argtypes = [vp]
    L.b3m_engine_last_error.restype = C.c_char_p
    L.b3m_engine_load_host.argtypes = [vp, vp, u64, C.c_int]
    L.b3m_engine_load_device.argtypes = [vp, vp, u64, C.c_int]
    L.b3m_engine_build.argtypes = [vp, C.POINTER(BuildParams)]
    L.b3m_engine_info.argtypes = [vp, C.POINTER(Info)]
    L.b3m_engine_fetch.argtypes = [vp, vp, vp, vp, vp]
    L.b3m_engine_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.b3m_engine_lf_bench.argtypes = [vp, u64, u64, C.POINTER(C.c_float), u64p]
    L.b3m_engine_sync.argtypes = [vp]
    L.b3m_engine_set_profile.argtypes = [vp, C.c_int]
    L.b3m_engine_kernel_times.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.b3m_engine_write_bwt.argtypes = [vp, C.c_char_p]
    L.b3m_engine_fetch_runs.argtypes = [vp, vp, vp, u64, u64p]
    L.b3m_engine_ssa_from_bwt.argtypes = [vp, vp, u64, vp, u64, u64, u64]
    L.b3m_engine_fetch_bwa.argtypes = [vp, vp, u64, u64p, u64p, u64p]
    L.b3m_engine_shard_build.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(BuildParams), vp, vp, vp, vp, vp, u64p]
    L.b3m_engine_shard_finish.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint32]
    L.b3m_engine_shard_adopt.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint32]
    L.b3m_engine_xshard_count.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(BuildParams), vp, C.POINTER(C.c_uint32)]
    L.b3m_engine_xshard_scatter.argtypes = [vp, vp, vp, vp]
    L.b3m_engine_xshard_finish.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64p]
    L.b3m_engine_xshard_stream_sa.argtypes = [vp, vp, vp]
    L.b3m_engine_xshard_sa_delivered.argtypes = [vp, C.POINTER(C.c_int)]
    L.b3m_dev_alloc.argtypes = [C.c_int, u64, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.b3m_dev_free.argtypes = [C.c_int, vp, C.c_char_p, C.c_size_t]
    L.b3m_ipc_export.argtypes = [C.c_int, vp, C.c_char_p, C.c_char_p, C.c_size_t]
    L.b3m_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.b3m_ipc_close.argtypes = [C.c_int, vp, C.c_char_p, C.c_size_t]
    L.b3m_dev_copy.argtypes = [C.c_int, vp, vp, u64, vp, C.c_char_p, C.c_size_t]
    L.b3m_engine_pack_bwa.argtypes = [vp, vp, u64, u64]
    L.b3m_multi_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.b3m_multi_destroy.argtypes = [vp]
    L.b3m_multi_destroy.restype = None
    L.b3m_multi_last_error.argtypes = [vp]
    L.b3m_multi_last_error.restype = C.c_char_p
    L.b3m_multi_load_host.argtypes = [vp, vp, u64, C.c_int]
    L.b3m_multi_build.argtypes = [vp, C.POINTER(BuildParams)]
    L.b3m_multi_engine.argtypes = [vp, C.c_int]
    L.b3m_multi_engine.restype = vp
    L.b3m_multi_stats.argtypes = [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.b3m_engine_shard_rows.argtypes = [vp, C.c_uint32, u64p]
    L.b3m_engine_pack_rows.argtypes = [vp, vp, u64, vp]
    L.b3m_engine_unpack_rows.argtypes = [vp, vp, u64, vp]
    L.b3m_bwt_length.argtypes = [C.c_char_p, u64p, C.c_char_p, C.c_size_t]
    L.b3m_bwt_decode.argtypes = [C.c_char_p, vp, u64, u64, C.c_char_p, C.c_size_t]
    L.b3m_bwt_encode_host.argtypes = [C.c_char_p, vp, u64, C.c_char_p, C.c_size_t]
    L.b3m_bwt_block_sym_histograms.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, u64, u64p, C.c_char_p, C.c_size_t]
    L.b3m_bwt_rank.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, u64, u64p, C.c_char_p, C.c_size_t]
    L.b3m_compact_write.argtypes = [C.c_char_p, C.c_uint, vp, u64, C.c_char_p, C.c_size_t]
    L.b3m_compact_info.argtypes = [C.c_char_p, C.POINTER(C.c_uint), u64p, C.c_char_p, C.c_size_t]
    L.b3m_compact_read.argtypes = [C.c_char_p, vp, u64, C.c_char_p, C.c_size_t]
    u32p = C.POINTER(C.c_uint32)
    L.b3m_engine_default_preisarate.argtypes = [vp, C.c_int, u64p]
    L.b3m_engine_blk_begin.argtypes = [vp, u64, u64, vp, vp, vp]
    L.b3m_engine_blk_build_range.argtypes = [vp, u64, u64, u64, vp, u32p]
    L.b3m_engine_blk_chains.argtypes = [vp, u64, u64p, u64p]
    L.b3m_engine_blk_zranks.argtypes = [vp, u64, u64, u64, u64, u64, vp]
    L.b3m_engine_blk_gap.argtypes = [vp, vp, u64, u64, C.c_uint32, u64, u64, u64, u64, u64, vp, vp, vp]
    L.b3m_engine_blk_merge.argtypes = [vp, vp, u64, C.c_uint32, vp, u64, C.c_uint32, u64, vp, vp, u32p]
    L.b3m_engine_blk_merge_samples.argtypes = [vp, u64, u64, u64, vp]
    L.b3m_engine_blk_finish.argtypes = [vp, vp, C.c_uint32, u64, u64, u64, u64, C.c_int, u64]
    if True:
        L.b3m_options_init.argtypes = [C.POINTER(Options)]
        L.b3m_options_init.restype = None
        L.b3m_compute_bwt.argtypes = [C.POINTER(Options), C.POINTER(Result), C.c_char_p, C.c_size_t]
        L.b3m_compute_ssa.argtypes = [C.c_char_p, u64, u64, C.c_char_p, C.c_int, u64, u64, u64, C.c_int, C.c_char_p,
                                      C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.b3m_to_bwa.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.b3m_lf_speed.argtypes = [C.c_char_p, u64, u64, u64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p, C.c_size_t]
        L.b3m_check_bwt.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, u64, C.c_int, C.c_int, C.POINTER(C.c_int), u64p, C.c_char_p, C.c_size_t]
    _lib = L
    return L
