"""Python mirror of the reference's Julia API, over the C ABI of libsmslu.so.

The reference (johnomotani/SharedMemSparseLU.jl, src/SharedMemSparseLU.jl) exposes
``ParallelSparseLU(A[, chunk_size])`` (src:64), ``lu!(F, A)`` (src:245), ``ldiv!(x, F, b)``
(src:286), ``lsolve!(F, x)`` (src:349), ``rsolve!(F, x)`` (src:374) and the fields
``F.m F.n F.L F.U F.p F.q F.Rs`` (src:45-51).  No Julia toolchain exists in this image, so the
host-side mirror used by the tests and the benchmark is this module (Julia's ``!`` becomes a
trailing underscore); the Julia shim that forwards the same calls through ``ccall`` lives in
``julia/SharedMemSparseLU.jl``.  All numeric work happens in hand-written sm_100a kernels
behind ``libsmslu.so``; nothing here computes.

Index conventions follow Python (0-based ``p``/``q``); the Julia shim passes ``index_base=1``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import DimensionMismatch, PivotThresholdError, SingularException, SmsluError  # noqa: F401

__all__ = ["ParallelSparseLU", "lu_", "ldiv_", "lsolve_", "rsolve_", "cleanup_ParallelSparseLU_",
           "allocate_shared", "pinned_empty", "comm_unique_id", "host_pivots", "DimensionMismatch", "SingularException",
           "PivotThresholdError", "SmsluError"]


def comm_unique_id() -> bytes:
    """128-byte NCCL id made on rank 0; broadcast it to the other ranks (torch.distributed, MPI, ...)
    and pass it to every rank's ``ParallelSparseLU(..., nranks=N, rank=r, comm_id=id)``."""
    buf = C.create_string_buffer(128)
    rc = _capi.lib().smslu_comm_unique_id(buf, 128)
    if rc != 0:
        raise _capi.SmsluError(rc, "smslu_comm_unique_id failed")
    return buf.raw


def host_pivots(A, pivot_tol=1.0e-3):
    """Host pivot search: (p, q, Rs) with a threshold-pivoted L*U == (Rs .* A)[p, q].

    What the Julia shim gets from UMFPACK's ``lu(A)`` (src:74) -- ordering, threshold partial pivoting with
    a preference for the diagonal, row scaling Rs = 1/sum_j|a_ij| -- comes here from SciPy's SuperLU (UMFPACK is
    not installed): COLAMD column ordering, diagonal preferred when it passes ``pivot_tol``.  Only the
    permutations are used; every factor entry is then computed on the GPU under that static pivot order."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    A = sp.csc_matrix(A)
    rs = np.asarray(abs(A).sum(axis=1)).ravel()
    Rs = np.where(rs > 0, 1.0 / np.where(rs > 0, rs, 1.0), 1.0)
    B = sp.csc_matrix(sp.diags(Rs) @ A)
    lu = spla.splu(B, permc_spec="COLAMD", diag_pivot_thresh=float(pivot_tol), options=dict(Equil=False))
    return np.argsort(lu.perm_r).astype(np.int64), np.argsort(lu.perm_c).astype(np.int64), Rs


def _ptr(a):
    """Raw address of a numpy array (host) or of anything with data_ptr() (torch CUDA tensor)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _len(a):
    return int(a.shape[0])


def _check_f64(a, name):
    if hasattr(a, "data_ptr"):   # torch tensor
        import torch
        if a.dtype != torch.float64 or not a.is_contiguous():
            raise TypeError("%s must be a contiguous float64 tensor" % name)
        return a
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not (a.flags.c_contiguous or a.flags.f_contiguous):
        raise TypeError("%s must be a contiguous float64 numpy array (or a torch CUDA tensor)" % name)
    return a


class _Pinned:
    """Owner of a page-locked host buffer exposed as a numpy array."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        rc = _capi.lib().smslu_host_alloc(C.byref(self.ptr), nbytes)
        if rc != 0:
            raise _capi.SmsluError(rc, "smslu_host_alloc failed")

    def __del__(self):
        try:
            if self.ptr:
                _capi.lib().smslu_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(n, dtype=np.float64):
    """numpy array backed by page-locked host memory (for x / b / nzval in timed loops)."""
    dt = np.dtype(dtype)
    owner = _Pinned(max(int(n), 1) * dt.itemsize)
    buf = (C.c_char * (max(int(n), 1) * dt.itemsize)).from_address(owner.ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    arr = arr.view(_PinnedArray)
    arr._owner = owner
    return arr


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


class ParallelSparseLU:
    """Factor object; mirrors the reference struct (src:43-62) and constructor (src:64-98).

    ``A`` is a scipy CSC matrix (Float64).  ``chunk_size`` is accepted for signature
    compatibility and ignored: the dense column-chunk layout (src:101-178) is replaced by the
    supernodal layout.  Extensions (keyword only): ``ordering`` ("auto", "natural", "given",
    "nd_graph", "nd_grid"), ``grid`` (nx, ny, nz), ``p``/``q``/``Rs`` to reproduce a given
    factorization contract ``L*U == (Rs .* A)[p, q]`` (what the Julia shim passes from
    UMFPACK's ``lu(A)``), ``scaling`` ("sum" = UMFPACK's default row scaling, "none").
    Multi-GPU (one process per GPU): ``nranks``, ``rank`` and the 128-byte ``comm_id`` from
    ``comm_unique_id()`` of rank 0; every call is then collective, every rank passes the full
    ``A`` / ``b`` and receives the full ``x``; ``F.L`` / ``F.U`` hold this rank's share of the values.
    """

    def __init__(self, A, chunk_size=None, *, pivots="auto", ordering="auto", grid=None, p=None, q=None, Rs=None,
                 scaling="sum", nd_leaf=None, relax=True, max_width=None, device=None,
                 nranks=1, rank=0, comm_id=None, pivot_tol=None):
        import scipy.sparse as sp
        if not sp.isspmatrix_csc(A):
            raise TypeError("A must be a scipy.sparse.csc_matrix (SparseMatrixCSC)")
        if A.shape[0] != A.shape[1]:
            raise DimensionMismatch(_capi.E_DIM, "matrix is not square")
        if pivots not in ("auto", "native", "host"):
            raise ValueError("pivots must be 'auto', 'native' or 'host'")
        # pivots: "native" = the library's nested-dissection ordering with diagonal pivots, never re-pivoted (a failed
        # threshold test raises PivotThresholdError); "host" = pivot search on the host (host_pivots) at construction and
        # again whenever lu_ fails the threshold test; "auto" = native first, host pivots when the test fails.
        self._pivots = pivots if (p is None and nranks == 1) else "native"
        self._ctor = dict(ordering=ordering, grid=grid, scaling=scaling, nd_leaf=nd_leaf, relax=relax, max_width=max_width,
                          device=device, nranks=nranks, rank=rank, comm_id=comm_id, pivot_tol=pivot_tol)
        self._h = None
        if self._pivots == "host":
            p, q, Rs = host_pivots(A, 1.0e-3 if pivot_tol in (None, 0) else abs(pivot_tol))
        try:
            self._setup(A, chunk_size, p, q, Rs)
        except (PivotThresholdError, SingularException):
            if self._pivots != "auto":
                raise
            self._repivot(A)

    def _repivot(self, A):
        """The static pivot order failed for these values: fresh host pivot search, re-analysis, refactorization
        (the reference's lu! lets UMFPACK re-pivot and re-chunks when the pattern of L/U changed, src:252-273)."""
        import scipy.sparse as sp
        A = sp.csc_matrix((A.data, self._rowval, self._colptr), shape=(self.n, self.n)) if not hasattr(A, "indptr") else A
        tol = self._ctor["pivot_tol"]
        try:
            p, q, Rs = host_pivots(A, 1.0e-3 if tol in (None, 0) else abs(tol))
        except RuntimeError as exc:              # SuperLU: "Factor is exactly singular"
            raise SingularException(_capi.E_PIVOT, "host pivot search: %s" % exc) from None
        self.close()
        self._pivots = "host"
        self._setup(A, self.chunk_size, p, q, Rs)

    def _setup(self, A, chunk_size, p, q, Rs):
        ordering, grid, scaling, nd_leaf, relax, max_width, device, nranks, rank, comm_id, pivot_tol = (
            self._ctor[k] for k in ("ordering", "grid", "scaling", "nd_leaf", "relax", "max_width", "device", "nranks",
                                    "rank", "comm_id", "pivot_tol"))
        L = _capi.lib()
        self.chunk_size = 8 if chunk_size is None else chunk_size   # kept, unused
        self.m = self.n = int(A.shape[0])
        self._colptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self._rowval = np.ascontiguousarray(A.indices, dtype=np.int64)
        opts = _capi.Options()
        L.smslu_options_default(C.byref(opts))
        if (p is None) != (q is None):
            raise ValueError("p and q must be given together")
        if p is not None:
            ordering = "given"
        opts.ordering = _capi.ORD[ordering]
        if grid is not None:
            g = list(grid) + [1] * (3 - len(grid))
            for d in range(3):
                opts.grid[d] = int(g[d])
        if nd_leaf is not None:
            opts.nd_leaf = int(nd_leaf)
        opts.relax = 1 if relax else 0
        if max_width is not None:
            opts.max_width = int(max_width)
        opts.scaling = _capi.SCALE[scaling]
        if device is not None:
            opts.device = int(device)
        opts.nranks, opts.rank = int(nranks), int(rank)
        if pivot_tol is not None:
            opts.pivot_tol = float(pivot_tol)
        self.nranks, self.rank = int(nranks), int(rank)
        self._h = C.c_void_p()
        rc = L.smslu_create(C.byref(self._h), self.n, _ptr(self._colptr), _ptr(self._rowval), 0, C.byref(opts))
        if rc != 0:
            raise _capi.SmsluError(rc, "smslu_create failed")
        pp = None if p is None else np.ascontiguousarray(p, dtype=np.int64)
        qq = None if q is None else np.ascontiguousarray(q, dtype=np.int64)
        _capi.check(self._h, L.smslu_analyze(self._h, _ptr(pp), _ptr(qq)))
        if self.nranks > 1 and comm_id is not None:      # collective: every rank of the partition
            _capi.check(self._h, L.smslu_comm_init(self._h, C.c_char_p(bytes(comm_id)), len(comm_id)))
        self._Rs_given = None if Rs is None else np.ascontiguousarray(Rs, dtype=np.float64)
        self._cache = {}
        self._constructed = False
        self._numeric(A)
        self._constructed = True
        # lu!(F, A) with new values: the row scaling is recomputed from them (UMFPACK's lu! does the same), unless the
        # caller fixed the whole contract (p, q, Rs) explicitly
        if self._pivots == "host":
            self._Rs_given = None

    # -- numeric (re)factorization -------------------------------------------------------------
    def _numeric(self, A):
        if A is None:   # the reference's `Nothing` arm (src:246) has no working method either
            raise TypeError("lu!(F, nothing) is not supported")
        if hasattr(A, "indptr"):
            if A.shape != (self.n, self.n):
                raise DimensionMismatch(_capi.E_DIM, "matrix size differs from the factor object")
            if A.indptr.shape != self._colptr.shape or A.indices.shape != self._rowval.shape or \
                    not (np.array_equal(A.indptr, self._colptr) and np.array_equal(A.indices, self._rowval)):
                raise _capi.SmsluError(_capi.E_PATTERN, "sparsity pattern differs from the analysed one")
            vals = np.ascontiguousarray(A.data, dtype=np.float64)
        else:           # raw nzval (numpy / pinned / torch CUDA tensor), same pattern
            vals = _check_f64(A, "nzval")
        self._cache = {}
        try:
            _capi.check(self._h, _capi.lib().smslu_refactor(self._h, _ptr(vals), _ptr(self._Rs_given)))
        except (PivotThresholdError, SingularException):
            # constructor: handled there.  lu!(F, A) with values that break the current pivots: re-pivot on the host
            if self._pivots == "native" or not getattr(self, "_constructed", False) or hasattr(vals, "data_ptr"):
                raise
            import scipy.sparse as sp
            self._repivot(sp.csc_matrix((np.asarray(vals), self._rowval, self._colptr), shape=(self.n, self.n)))

    # -- fields ----------------------------------------------------------------------------------
    def _factors(self):
        if "L" not in self._cache:
            import scipy.sparse as sp
            L = _capi.lib()
            n = self.n
            nl, nu = C.c_int64(), C.c_int64()
            _capi.check(self._h, L.smslu_get_nnz(self._h, C.byref(nl), C.byref(nu)))
            lp = np.zeros(n + 1, np.int64); li = np.zeros(nl.value, np.int64); lx = np.zeros(nl.value)
            up = np.zeros(n + 1, np.int64); ui = np.zeros(nu.value, np.int64); ux = np.zeros(nu.value)
            p = np.zeros(n, np.int64); q = np.zeros(n, np.int64); Rs = np.zeros(n)
            _capi.check(self._h, L.smslu_get_factors(self._h, _ptr(lp), _ptr(li), _ptr(lx), _ptr(up), _ptr(ui),
                                                    _ptr(ux), _ptr(p), _ptr(q), _ptr(Rs), 0))
            self._cache.update(L=sp.csc_matrix((lx, li, lp), shape=(n, n)),
                               U=sp.csc_matrix((ux, ui, up), shape=(n, n)), p=p, q=q, Rs=Rs)
        return self._cache

    @property
    def L(self):
        return self._factors()["L"]

    @property
    def U(self):
        return self._factors()["U"]

    @property
    def p(self):
        if "p" not in self._cache:
            p = np.zeros(self.n, np.int64); q = np.zeros(self.n, np.int64)
            _capi.check(self._h, _capi.lib().smslu_get_factors(self._h, None, None, None, None, None, None,
                                                              _ptr(p), _ptr(q), None, 0))
            self._cache.update(p=p, q=q)
        return self._cache["p"]

    @property
    def q(self):
        self.p
        return self._cache["q"]

    @property
    def Rs(self):
        return self._factors()["Rs"]

    # -- stream-ordered device-only variants (kernel-only timing, CUDA-array callers) -------------
    def set_stream(self, cuda_stream):
        """Run all of this handle's work on the caller's stream (int / torch.cuda.Stream)."""
        ptr = getattr(cuda_stream, "cuda_stream", cuda_stream)
        _capi.check(self._h, _capi.lib().smslu_set_stream(self._h, C.c_void_p(int(ptr))))

    def set_profile(self, on=True):
        _capi.check(self._h, _capi.lib().smslu_set_profile(self._h, 1 if on else 0))

    def refactor_async(self, nzval_dev, Rs_dev=None):
        self._cache = {}
        _capi.check(self._h, _capi.lib().smslu_refactor_async(self._h, _ptr(nzval_dev), _ptr(Rs_dev)))

    def solve_async(self, x_dev, b_dev):
        _capi.check(self._h, _capi.lib().smslu_solve_async(self._h, _ptr(x_dev), _ptr(b_dev)))

    def sync(self):
        _capi.check(self._h, _capi.lib().smslu_sync(self._h))

    def stats(self):
        st = _capi.Stats()
        _capi.check(self._h, _capi.lib().smslu_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def symbolic(self):
        """Supernodal layout read-back (tests)."""
        st = self.stats()
        nsn, sr, n = st["n_supernodes"], st["sum_rows"], self.n
        out = dict(sn_start=np.zeros(nsn + 1, np.int64), rows_ptr=np.zeros(nsn + 1, np.int64),
                   rows=np.zeros(sr, np.int64), sn_parent=np.zeros(nsn, np.int64),
                   sn_level=np.zeros(nsn, np.int64), etree_parent=np.zeros(n, np.int64),
                   colcount=np.zeros(n, np.int64))
        _capi.check(self._h, _capi.lib().smslu_get_symbolic(self._h, *[_ptr(out[k]) for k in (
            "sn_start", "rows_ptr", "rows", "sn_parent", "sn_level", "etree_parent", "colcount")]))
        return out

    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().smslu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _SymbolicOnly(ParallelSparseLU):
    """Analysis without the numeric step (CPU-only tests of the host logic)."""

    def _numeric(self, A):
        return None


def lu_(F: ParallelSparseLU, A):
    """``lu!(F, A)`` (src:245-279): numeric refactorization with the analysed pattern.  Returns None."""
    F._numeric(A)
    return None


def ldiv_(x, F: ParallelSparseLU, b):
    """``ldiv!(x, F, b)`` (src:286-342): solve A x = b, overwrite and return x; b is untouched.
    x, b: float64 vectors (numpy or torch CUDA); a 2-D Fortran-ordered block solves many RHS."""
    _check_f64(x, "x"); _check_f64(b, "b")
    nrhs = 1 if len(x.shape) == 1 else int(x.shape[1])
    nb = 1 if len(b.shape) == 1 else int(b.shape[1])
    if nrhs != nb:
        raise DimensionMismatch(_capi.E_DIM, "x and b have different numbers of columns")
    if nrhs > 1 and not (hasattr(x, "data_ptr") or (x.flags.f_contiguous and b.flags.f_contiguous)):
        raise TypeError("multi-RHS blocks must be Fortran-ordered")
    _capi.check(F._h, _capi.lib().smslu_solve(F._h, _ptr(x), _len(x), _ptr(b), _len(b), nrhs,
                                             _len(x), _len(b)))
    return x


def lsolve_(F: ParallelSparseLU, x):
    """``lsolve!(F, x)`` (src:349-367): x <- L^{-1} x in place.  Returns None."""
    _check_f64(x, "x")
    nrhs = 1 if len(x.shape) == 1 else int(x.shape[1])
    _capi.check(F._h, _capi.lib().smslu_lsolve(F._h, _ptr(x), _len(x), nrhs, _len(x)))
    return None


def rsolve_(F: ParallelSparseLU, x):
    """``rsolve!(F, x)`` (src:374-392): x <- U^{-1} x in place.  Returns None."""
    _check_f64(x, "x")
    nrhs = 1 if len(x.shape) == 1 else int(x.shape[1])
    _capi.check(F._h, _capi.lib().smslu_rsolve(F._h, _ptr(x), _len(x), nrhs, _len(x)))
    return None


def cleanup_ParallelSparseLU_(F: ParallelSparseLU):
    """``cleanup_ParallelSparseLU!`` is exported by the reference but never defined (src:31);
    here it releases the handle's device memory."""
    F.close()
    return None


def allocate_shared(*args, **kwargs):
    """Exported by the reference but never defined (src:31); kept as a documented no-op."""
    _capi.lib().smslu_allocate_shared()
    return None
