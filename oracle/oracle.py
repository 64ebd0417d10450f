"""ctypes front-end of the CPU oracle (oracle/ref_lu.c, oracle/ref_chunks.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs -- never from the product package.

PARITY UNPINNED (see the header of ref_lu.c): the reference's factorization is UMFPACK's,
which is not available here, and the reference's tests hold no golden vectors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = [os.path.join(_HERE, f) for f in ("ref_lu.c", "ref_chunks.c")]

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc if it is missing or older than its sources."""
    stale = force or not os.path.exists(_LIB) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB) for s in _SRC)
    if stale:
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        tmp = _LIB + ".%d.tmp" % os.getpid()
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", tmp] + _SRC + ["-lm"])
        os.replace(tmp, _LIB)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_lu_factor.restype = C.c_void_p
        L.oracle_lu_factor.argtypes = [C.c_longlong, _i64p, _i64p, _f64p, C.c_void_p, _i64p,
                                       C.c_void_p, C.c_int, C.c_double]
        L.oracle_lu_free.argtypes = [C.c_void_p]
        for f in ("oracle_lu_n", "oracle_lu_bad_col", "oracle_lu_nnzL", "oracle_lu_nnzU"):
            getattr(L, f).restype = C.c_longlong
            getattr(L, f).argtypes = [C.c_void_p]
        L.oracle_lu_flops.restype = C.c_double
        L.oracle_lu_flops.argtypes = [C.c_void_p]
        L.oracle_lu_export.argtypes = [C.c_void_p, _i64p, _i64p, _f64p, _i64p, _i64p, _f64p,
                                       _i64p, _i64p, _f64p]
        L.oracle_row_scale_sum.argtypes = [C.c_longlong, _i64p, _i64p, _f64p, _f64p]
        L.oracle_csc_lsolve.argtypes = [C.c_longlong, _i64p, _i64p, _f64p, _f64p]
        L.oracle_csc_usolve.argtypes = [C.c_longlong, _i64p, _i64p, _f64p, _f64p]
        L.oracle_lu_solve.argtypes = [C.c_longlong, _i64p, _i64p, _f64p, _i64p, _i64p, _f64p,
                                      _i64p, _i64p, _f64p, _f64p, _f64p, _f64p]
        L.ref_chunks_build.restype = C.c_void_p
        L.ref_chunks_build.argtypes = [C.c_longlong, C.c_longlong, _i64p, _i64p, _f64p,
                                       _i64p, _i64p, _f64p]
        L.ref_chunks_free.argtypes = [C.c_void_p]
        L.ref_chunks_bytes.restype = C.c_double
        L.ref_chunks_bytes.argtypes = [C.c_void_p]
        L.ref_chunks_predict_bytes.restype = C.c_double
        L.ref_chunks_predict_bytes.argtypes = [C.c_longlong, C.c_longlong, _i64p, _i64p, _i64p, _i64p]
        L.ref_chunks_total.restype = C.c_longlong
        L.ref_chunks_total.argtypes = [C.c_void_p]
        L.ref_chunks_ranges.argtypes = [C.c_void_p] + [_i64p] * 8
        L.ref_chunks_lsolve.argtypes = [C.c_void_p, _f64p]
        L.ref_chunks_rsolve.argtypes = [C.c_void_p, _f64p]
        L.ref_chunks_ldiv.restype = C.c_int
        L.ref_chunks_ldiv.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, _i64p, _i64p, _f64p,
                                      _f64p, _f64p, _f64p]
        _lib = L
    return _lib


def _csc(A):
    import scipy.sparse as sp
    A = sp.csc_matrix(A)
    A.sort_indices()
    return (A.shape[0], np.ascontiguousarray(A.indptr, np.int64),
            np.ascontiguousarray(A.indices, np.int64), np.ascontiguousarray(A.data, np.float64))


def row_scale_sum(A) -> np.ndarray:
    n, Ap, Ai, Ax = _csc(A)
    Rs = np.empty(n)
    lib().oracle_row_scale_sum(n, Ap, Ai, Ax, Rs)
    return Rs


def csc_lsolve(L, b):
    """`L \\ b` for a lower-triangular CSC matrix with its diagonal stored (test/runtests.jl:51,70)."""
    n, Lp, Li, Lx = _csc(L)
    x = np.array(b, np.float64, copy=True)
    lib().oracle_csc_lsolve(n, Lp, Li, Lx, x)
    return x


def csc_usolve(U, b):
    """`U \\ b` for an upper-triangular CSC matrix (test/runtests.jl:86,104)."""
    n, Up, Ui, Ux = _csc(U)
    x = np.array(b, np.float64, copy=True)
    lib().oracle_csc_usolve(n, Up, Ui, Ux, x)
    return x


class OracleLU:
    """L*U == (Rs .* A)[p, q] (0-based p, q), restating the contract at reference src:305-316.

    ``p is None`` => threshold partial pivoting with diagonal preference (stand-in for the
    pivot search UMFPACK would do); otherwise the given pivot rows are obeyed (static).
    """

    def __init__(self, A, q=None, p=None, Rs=None, diag_tol: float = 1.0e-3):
        import scipy.sparse as sp
        n, Ap, Ai, Ax = _csc(A)
        self.n = n
        q = np.arange(n, dtype=np.int64) if q is None else np.ascontiguousarray(q, np.int64)
        mode = 1 if p is None else 0
        pp = None if p is None else np.ascontiguousarray(p, np.int64)
        rs = None if Rs is None else np.ascontiguousarray(Rs, np.float64)
        h = lib().oracle_lu_factor(n, Ap, Ai, Ax,
                                   None if pp is None else pp.ctypes.data_as(C.c_void_p), q,
                                   None if rs is None else rs.ctypes.data_as(C.c_void_p),
                                   mode, float(diag_tol))
        if not h:
            raise MemoryError("oracle_lu_factor")
        try:
            self.bad_col = int(lib().oracle_lu_bad_col(h))
            self.flops = float(lib().oracle_lu_flops(h))
            nl, nu = int(lib().oracle_lu_nnzL(h)), int(lib().oracle_lu_nnzU(h))
            Lp = np.zeros(n + 1, np.int64); Li = np.zeros(nl, np.int64); Lx = np.zeros(nl)
            Up = np.zeros(n + 1, np.int64); Ui = np.zeros(nu, np.int64); Ux = np.zeros(nu)
            self.p = np.zeros(n, np.int64); self.q = np.zeros(n, np.int64); self.Rs = np.zeros(n)
            lib().oracle_lu_export(h, Lp, Li, Lx, Up, Ui, Ux, self.p, self.q, self.Rs)
        finally:
            lib().oracle_lu_free(h)
        self.Lp, self.Li, self.Lx, self.Up, self.Ui, self.Ux = Lp, Li, Lx, Up, Ui, Ux
        self.L = sp.csc_matrix((Lx, Li, Lp), shape=(n, n))
        self.U = sp.csc_matrix((Ux, Ui, Up), shape=(n, n))

    def lsolve(self, x):
        x = np.array(x, np.float64, copy=True)
        lib().oracle_csc_lsolve(self.n, self.Lp, self.Li, self.Lx, x)
        return x

    def usolve(self, x):
        x = np.array(x, np.float64, copy=True)
        lib().oracle_csc_usolve(self.n, self.Up, self.Ui, self.Ux, x)
        return x

    def solve(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty(self.n); w = np.empty(self.n)
        lib().oracle_lu_solve(self.n, self.Lp, self.Li, self.Lx, self.Up, self.Ui, self.Ux,
                              self.p, self.q, self.Rs, b, x, w)
        return x


class RefChunks:
    """The reference's dense column-chunk solver (src:101-243, 349-392) on given CSC factors."""

    def __init__(self, L, U, chunk_size=None):
        n, Lp, Li, Lx = _csc(L)
        _, Up, Ui, Ux = _csc(U)
        self.n = n
        self._h = lib().ref_chunks_build(n, -1 if chunk_size is None else int(chunk_size),
                                         Lp, Li, Lx, Up, Ui, Ux)
        if not self._h:
            raise MemoryError("ref_chunks_build")
        self.total_chunks = int(lib().ref_chunks_total(self._h))
        self.bytes = float(lib().ref_chunks_bytes(self._h))

    @staticmethod
    def predict_bytes(L, U, chunk_size=8) -> float:
        n, Lp, Li, _ = _csc(L)
        _, Up, Ui, _ = _csc(U)
        return float(lib().ref_chunks_predict_bytes(n, chunk_size, Lp, Li, Up, Ui))

    def ranges(self):
        T = self.total_chunks
        arrs = [np.zeros(T, np.int64) for _ in range(8)]
        lib().ref_chunks_ranges(self._h, *arrs)
        keys = ("lc0", "lc1", "lr0", "lr1", "uc0", "uc1", "ur0", "ur1")
        return dict(zip(keys, arrs))

    def lsolve(self, x):
        x = np.array(x, np.float64, copy=True)
        lib().ref_chunks_lsolve(self._h, x)
        return x

    def rsolve(self, x):
        x = np.array(x, np.float64, copy=True)
        lib().ref_chunks_rsolve(self._h, x)
        return x

    def ldiv(self, p, q, Rs, b, x=None):
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty(self.n) if x is None else x
        w = np.empty(self.n)
        rc = lib().ref_chunks_ldiv(self._h, x.shape[0], b.shape[0], np.ascontiguousarray(p, np.int64),
                                   np.ascontiguousarray(q, np.int64), np.ascontiguousarray(Rs, np.float64),
                                   b, x, w)
        if rc != 0:
            raise ValueError("DimensionMismatch")   # reference src:288-290
        return x

    def close(self):
        if self._h:
            lib().ref_chunks_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
