"""ctypes binding of include/beluga_b200.h.  There is no CPU fallback: if the shared library is
missing or no CUDA device is present, loading / blg_create raise."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libbeluga_b200.so")

_lib = None

# every symbol include/beluga_b200.h declares
SYMBOLS = [
    "blg_create", "blg_destroy", "blg_last_error", "blg_set_profiles", "blg_set_profiles_columns", "blg_set_topology", "blg_set_params",
    "blg_rebuild", "blg_logpdf_full", "blg_logpdf_from", "blg_logpdf_root", "blg_logpdf_full_async",
    "blg_logpdf_from_async", "blg_logpdf_root_async", "blg_gradient", "blg_gradient_cr", "blg_comm_unique_id", "blg_comm_init", "blg_comm_size",
    "blg_commit", "blg_revert", "blg_extend", "blg_shrink", "blg_get_column", "blg_get_family_logpdf",
    "blg_get_eps", "blg_get_W", "blg_ncols", "blg_npatterns", "blg_last_algorithmic", "blg_sync",
    "blg_set_timing", "blg_last_timing", "blg_family_order", "blg_fp64_peak", "blg_kernel_runs", "blg_pattern_stats", "blg_graph_stats", "blg_site_patterns", "blg_site_patterns_u16", "blg_rand_profiles",
]


class BelugaError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise BelugaError(
            "libbeluga_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback." % SO
        )
    # the library binds NCCL at run time (multi-GPU only).  In a Python process PyTorch may be imported later and brings its
    # own libnccl.so.2: point the library at that copy, so that one process never holds two different NCCLs
    if "BLG_NCCL_LIB" not in os.environ:
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
                cand = os.path.join(base, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["BLG_NCCL_LIB"] = cand
                    break
        except Exception:
            pass
    L = C.CDLL(SO)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    i32p, i64p, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint8)
    L.blg_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.blg_destroy.argtypes = [vp]
    L.blg_last_error.argtypes = [vp]
    L.blg_last_error.restype = C.c_char_p
    L.blg_set_profiles.argtypes = [vp, i32p, i64p, C.c_int, C.c_int]
    L.blg_set_profiles_columns.argtypes = [vp, i32p, i64p, C.c_int, C.c_int]
    L.blg_set_topology.argtypes = [vp, C.c_int, i32p, u8p, i32p]
    L.blg_set_params.argtypes = [vp, C.c_int, dp, dp, dp, dp, C.c_double]
    L.blg_rebuild.argtypes = [vp, i32p, C.c_int]
    L.blg_logpdf_full.argtypes = [vp, dp]
    L.blg_logpdf_from.argtypes = [vp, C.c_int, dp]
    L.blg_logpdf_root.argtypes = [vp, dp]
    L.blg_logpdf_full_async.argtypes = [vp, vp]
    L.blg_logpdf_from_async.argtypes = [vp, C.c_int, vp]
    L.blg_logpdf_root_async.argtypes = [vp, vp]
    L.blg_gradient.argtypes = [vp, dp, C.c_int, dp]
    L.blg_gradient_cr.argtypes = [vp, dp, dp]
    L.blg_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.blg_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.blg_comm_size.argtypes = [vp]
    for f in ("blg_commit", "blg_revert", "blg_sync", "blg_npatterns"):
        getattr(L, f).argtypes = [vp]
    L.blg_extend.argtypes = [vp, C.c_int]
    L.blg_family_order.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.blg_shrink.argtypes = [vp, C.c_int]
    L.blg_get_column.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, C.c_int, ip]
    L.blg_get_family_logpdf.argtypes = [vp, dp, C.c_int]
    L.blg_get_eps.argtypes = [vp, dp, dp, C.c_int]
    L.blg_get_W.argtypes = [vp, C.c_int, dp, C.c_int, ip]
    L.blg_ncols.argtypes = [vp, C.c_int]
    L.blg_last_algorithmic.argtypes = [vp, dp, dp, ip]
    L.blg_kernel_runs.argtypes = [vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint), ip, ip]
    L.blg_pattern_stats.argtypes = [vp, dp, dp, dp]
    L.blg_graph_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.blg_rand_profiles.argtypes = [vp, C.c_int, C.c_ulonglong, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), ip, C.POINTER(C.c_uint16), C.POINTER(C.c_longlong)]
    L.blg_site_patterns.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64), ip]
    L.blg_site_patterns_u16.argtypes = [C.c_int, C.POINTER(C.c_uint16), C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64), ip]
    L.blg_set_timing.argtypes = [vp, C.c_int]
    L.blg_fp64_peak.argtypes = [vp, dp]
    L.blg_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for s in SYMBOLS:
        if s != "blg_last_error":
            getattr(L, s).restype = C.c_int
    _lib = L
    return L
