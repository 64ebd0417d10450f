"""ctypes binding of libsmslu.so (include/smslu.h).  Thin: every function maps 1:1 to a C entry
point.  There is no Python or CPU implementation of the numeric path behind it -- if the shared
library is missing, loading fails loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMSLU_LIB") or os.path.join(HERE, "libsmslu.so")   # SMSLU_LIB: A/B testing of builds

OK, E_DIM, E_PIVOT, E_PATTERN, E_ARG, E_CUDA, E_OOM, E_INTERNAL, E_NCCL, E_REPIVOT = 0, -1, -2, -3, -4, -5, -6, -7, -8, -9
ORD = {"auto": 0, "natural": 1, "given": 2, "nd_graph": 3, "nd_grid": 4}
SCALE = {"none": 0, "sum": 1}

EXPORTS = [
    "smslu_options_default", "smslu_create", "smslu_analyze", "smslu_refactor", "smslu_solve",
    "smslu_lsolve", "smslu_rsolve", "smslu_get_nnz", "smslu_get_factors", "smslu_get_stats",
    "smslu_last_error", "smslu_get_symbolic", "smslu_destroy", "smslu_allocate_shared",
    "smslu_host_alloc", "smslu_host_free", "smslu_version", "smslu_refactor_async",
    "smslu_solve_async", "smslu_sync", "smslu_set_stream", "smslu_set_profile",
    "smslu_comm_unique_id", "smslu_comm_init", "smslu_debug_trace",
]
KERNEL_KINDS = ["rowscale", "scatter", "zero_cb", "extend_add", "front_small", "panel", "gemm_cb",
                "permute_scale", "fwd", "bwd", "unpermute", "fwd_small", "bwd_small", "allreduce"]


class Options(C.Structure):
    _fields_ = [("ordering", C.c_int32), ("grid", C.c_int32 * 3), ("nd_leaf", C.c_int32),
                ("relax", C.c_int32), ("max_width", C.c_int32), ("scaling", C.c_int32),
                ("device", C.c_int32), ("reserved0", C.c_int32), ("nranks", C.c_int32), ("rank", C.c_int32),
                ("pivot_tol", C.c_double), ("reserved", C.c_int32 * 4)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n", "nnz_a", "nnz_l_exact", "nnz_u_exact", "nnz_l_stored", "nnz_u_stored", "n_supernodes",
        "n_levels", "max_front", "max_pivot_block", "max_children", "sum_rows", "lu_pool_doubles",
        "cb_pool_doubles")] + [(k, C.c_double) for k in (
            "flops_exact", "flops_stored", "ms_analyze", "ms_upload", "ms_refactor", "ms_solve",
            "ms_refactor_h2d", "ms_solve_h2d", "ms_solve_d2h")] + [(k, C.c_int64) for k in (
                "launches_refactor", "launches_solve", "n_refactor", "n_solve", "bad_pivot_col")] + [
        ("ms_kernel", C.c_double * 16), ("launches_kernel", C.c_int64 * 16)] + [(k, C.c_int64) for k in (
            "n_top_supernodes", "n_local_supernodes", "allreduce_doubles_refactor", "allreduce_doubles_solve",
            "threshold_col")] + [("reserved", C.c_int64 * 3)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("reserved", "ms_kernel", "launches_kernel")}
        d["ms_kernel"] = {k: self.ms_kernel[i] for i, k in enumerate(KERNEL_KINDS)}
        d["launches_kernel"] = {k: self.launches_kernel[i] for i, k in enumerate(KERNEL_KINDS)}
        return d


class SmsluError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("smslu error %d: %s" % (code, msg))
        self.code = code


class DimensionMismatch(SmsluError, ValueError):
    """SMSLU_E_DIM -- the reference throws DimensionMismatch (src:288-290)."""


class SingularException(SmsluError, ArithmeticError):
    """SMSLU_E_PIVOT -- zero pivot under the static pivot order."""


class PivotThresholdError(SmsluError):
    """SMSLU_E_REPIVOT -- the factorization finished, but a multiplier exceeded 1/pivot_tol: the static pivot order
    fails the threshold test for these values (UMFPACK's lu! would re-pivot here, src:245-279)."""


_lib = None


def lib():
    """Load libsmslu.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libsmslu.so is not built (%s). Run `python __graft_entry__.py` or "
                "`python sharedmemsparselu.jl_b200/build.py`; there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, dp = C.c_void_p, C.c_int64, C.c_int32, C.c_void_p
        L.smslu_options_default.argtypes = [C.POINTER(Options)]
        L.smslu_create.argtypes = [C.POINTER(vp), i64, vp, vp, i32, C.POINTER(Options)]
        L.smslu_analyze.argtypes = [vp, vp, vp]
        L.smslu_refactor.argtypes = [vp, dp, dp]
        L.smslu_solve.argtypes = [vp, dp, i64, dp, i64, i64, i64, i64]
        L.smslu_lsolve.argtypes = [vp, dp, i64, i64, i64]
        L.smslu_rsolve.argtypes = [vp, dp, i64, i64, i64]
        L.smslu_get_nnz.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.smslu_get_factors.argtypes = [vp] + [vp] * 9 + [i32]
        L.smslu_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.smslu_last_error.argtypes = [vp]
        L.smslu_last_error.restype = C.c_char_p
        L.smslu_get_symbolic.argtypes = [vp] + [vp] * 7
        L.smslu_destroy.argtypes = [vp]
        L.smslu_host_alloc.argtypes = [C.POINTER(vp), i64]
        L.smslu_host_free.argtypes = [vp]
        L.smslu_refactor_async.argtypes = [vp, dp, dp]
        L.smslu_solve_async.argtypes = [vp, dp, dp]
        L.smslu_sync.argtypes = [vp]
        L.smslu_set_stream.argtypes = [vp, vp]
        L.smslu_set_profile.argtypes = [vp, i32]
        L.smslu_debug_trace.argtypes = [vp]
        L.smslu_comm_unique_id.argtypes = [vp, i64]
        L.smslu_comm_init.argtypes = [vp, vp, i64]
        for name in EXPORTS:
            if name != "smslu_last_error":
                getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def check(h, rc):
    if rc == OK:
        return
    msg = lib().smslu_last_error(h).decode() if h else ""
    if rc == E_DIM:
        raise DimensionMismatch(rc, msg)
    if rc == E_PIVOT:
        raise SingularException(rc, msg)
    if rc == E_REPIVOT:
        raise PivotThresholdError(rc, msg)
    raise SmsluError(rc, msg)
