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


# ----------------------------------------------------------------------------------------------
# Threaded CPU multifrontal port (oracle/ref_mf.cpp): the CPU side of bench.py's comparison.
_MF_LIB = os.path.join(_HERE, "_build", "libref_mf.so")
_MF_SRC = [os.path.join(_HERE, "ref_mf.cpp"),
           os.path.join(os.path.dirname(_HERE), "sharedmemsparselu.jl_b200", "csrc", "symbolic.cpp")]
_MF_HDR = [os.path.join(os.path.dirname(_HERE), "sharedmemsparselu.jl_b200", "csrc", "symbolic.hpp")]


# OpenMP threads that spin after a parallel region starve the BLAS threads of the next call (measured: 13x slower)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")


def build_mf(force: bool = False) -> str:
    """g++ -fopenmp: ref_mf.cpp + the host-only analysis source csrc/symbolic.cpp (no kernels, no libsmslu.so)."""
    stale = force or not os.path.exists(_MF_LIB) or any(
        os.path.getmtime(s) > os.path.getmtime(_MF_LIB) for s in _MF_SRC + _MF_HDR)
    if stale:
        os.makedirs(os.path.dirname(_MF_LIB), exist_ok=True)
        tmp = _MF_LIB + ".%d.tmp" % os.getpid()
        subprocess.check_call(["g++", "-O3", "-fopenmp", "-pthread", "-fPIC", "-std=c++17", "-shared",
                               "-o", tmp] + _MF_SRC)
        os.replace(tmp, _MF_LIB)
    return _MF_LIB


_mf = None


def _blas_ptr(name):
    """Address of SciPy's bundled OpenBLAS routine (scipy.linalg.cython_blas capsule)."""
    import scipy.linalg.cython_blas as cb
    cap = cb.__pyx_capi__[name]
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap))


def _blas_set_threads_ptr():
    """openblas_set_num_threads of the OpenBLAS behind scipy.linalg (so that the fronts handled one per
    OpenMP thread call BLAS on the calling thread only); None if it cannot be found."""
    try:
        import scipy.linalg  # noqa: F401  (loads the library)
        from threadpoolctl import threadpool_info
        for info in threadpool_info():
            if info.get("internal_api") != "openblas" or "scipy.libs" not in info.get("filepath", ""):
                continue
            L = C.CDLL(info["filepath"])
            for nm in ("scipy_openblas_set_num_threads", "openblas_set_num_threads"):
                if hasattr(L, nm):
                    return C.cast(getattr(L, nm), C.c_void_p)
    except Exception:
        pass
    return None


def mf_lib():
    global _mf
    if _mf is None:
        L = C.CDLL(build_mf())
        L.mf_create.restype = C.c_void_p
        L.mf_create.argtypes = [C.c_longlong, _i64p, _i64p, C.c_int, C.c_void_p, C.c_int]
        L.mf_set_blas.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mf_free.argtypes = [C.c_void_p]
        L.mf_info.argtypes = [C.c_void_p, _i64p, _f64p]
        L.mf_perm.argtypes = [C.c_void_p, _i64p, _i64p]
        L.mf_factor.restype = C.c_longlong
        L.mf_factor.argtypes = [C.c_void_p, _f64p, C.c_void_p, _f64p]
        L.mf_solve.argtypes = [C.c_void_p, _f64p, _f64p]
        L.mf_times.argtypes = [C.c_void_p, _f64p]
        L.mf_get_factors.argtypes = [C.c_void_p, _i64p, _i64p, _f64p, _i64p, _i64p, _f64p, _f64p]
        _mf = L
    return _mf


class RefMF:
    """CPU multifrontal LU (static diagonal pivots under the same nested-dissection ordering and
    supernodes as the GPU path, BLAS-3 fronts, all host cores).  ``lu_(Ax)`` = numeric
    refactorization with the analysis reused (reference src:245-279), ``ldiv(b)`` = src:286-342."""

    def __init__(self, A, grid=None, threads=0, blas=True):
        n, Ap, Ai, Ax = _csc(A)
        self.n, self._Ap, self._Ai = n, Ap, Ai
        g = None if grid is None else np.array(list(grid) + [1] * (3 - len(grid)), dtype=np.int32)
        self._h = mf_lib().mf_create(n, Ap, Ai, 0, None if g is None else g.ctypes.data_as(C.c_void_p), int(threads))
        if not self._h:
            raise RuntimeError("mf_create failed")
        if blas:
            mf_lib().mf_set_blas(self._h, _blas_ptr("dgemm"), _blas_ptr("dtrsm"), _blas_set_threads_ptr())
        info = np.zeros(8, np.int64); fl = np.zeros(1)
        mf_lib().mf_info(self._h, info, fl)
        self.info = dict(zip(("n", "nsn", "nlevels", "nnzL", "lu_size", "threads", "big_fronts", "max_front"), map(int, info)))
        self.flops = float(fl[0])
        self.growth = 0.0

    def lu_(self, Ax, Rs=None):
        Ax = np.ascontiguousarray(Ax, np.float64)
        rs = None if Rs is None else np.ascontiguousarray(Rs, np.float64)
        g = np.zeros(1)
        bad = mf_lib().mf_factor(self._h, Ax, None if rs is None else rs.ctypes.data_as(C.c_void_p), g)
        self.growth = float(g[0])
        return int(bad)

    def ldiv(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty(self.n)
        mf_lib().mf_solve(self._h, b, x)
        return x

    def times(self):
        t = np.zeros(2)
        mf_lib().mf_times(self._h, t)
        return float(t[0]), float(t[1])

    def perm(self):
        p = np.zeros(self.n, np.int64); q = np.zeros(self.n, np.int64)
        mf_lib().mf_perm(self._h, p, q)
        return p, q

    def factors(self):
        import scipy.sparse as sp
        n, nl = self.n, self.info["nnzL"]
        Lp = np.zeros(n + 1, np.int64); Li = np.zeros(nl, np.int64); Lx = np.zeros(nl)
        Up = np.zeros(n + 1, np.int64); Ui = np.zeros(nl, np.int64); Ux = np.zeros(nl)
        Rs = np.zeros(n)
        mf_lib().mf_get_factors(self._h, Lp, Li, Lx, Up, Ui, Ux, Rs)
        return sp.csc_matrix((Lx, Li, Lp), shape=(n, n)), sp.csc_matrix((Ux, Ui, Up), shape=(n, n)), Rs

    def close(self):
        if self._h:
            mf_lib().mf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
