import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def smslu():
    """The product package (builds libsmslu.so on demand)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_smslu_build", os.path.join(ROOT, "sharedmemsparselu.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    import smslu as m
    return m


@pytest.fixture(scope="session")
def W(smslu):
    from sharedmemsparselu_jl_b200 import workloads
    return workloads


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.build()
    return oracle


class HostExec:
    """ctypes wrapper of tests/hostexec.cpp (CPU walk over the device data structures)."""

    def __init__(self):
        src = [os.path.join(ROOT, "tests", "hostexec.cpp"),
               os.path.join(ROOT, "sharedmemsparselu.jl_b200", "csrc", "symbolic.cpp")]
        hdr = os.path.join(ROOT, "sharedmemsparselu.jl_b200", "csrc", "symbolic.hpp")
        out = os.path.join(ROOT, "tests", "_build", "libhostexec.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in src + [hdr]):
            tmp = out + ".%d.tmp" % os.getpid()
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", tmp] + src)
            os.replace(tmp, out)
        L = C.CDLL(out)
        i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.hx_create.restype = C.c_void_p
        L.hx_create.argtypes = [C.c_longlong, i64p, i64p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p]
        L.hx_create2.restype = C.c_void_p
        L.hx_create2.argtypes = L.hx_create.argtypes + [C.c_int]
        L.hx_owner.argtypes = [C.c_void_p, i64p]
        L.hx_buffer.restype = C.c_void_p
        L.hx_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
        L.hx_factor_phase.restype = C.c_longlong
        L.hx_factor_phase.argtypes = [C.c_void_p, f64p, C.c_void_p, C.c_int, C.c_int]
        L.hx_top_levels.restype = C.c_int
        L.hx_top_levels.argtypes = [C.c_void_p, i64p]
        L.hx_top_segments.restype = C.c_int
        L.hx_top_segments.argtypes = [C.c_void_p, C.c_int, C.c_int, i64p]
        L.hx_top_panels.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hx_top_update.restype = C.c_longlong
        L.hx_top_update.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hx_lsolve_phase.argtypes = [C.c_void_p, f64p, C.c_int, C.c_int]
        L.hx_rsolve_phase.argtypes = [C.c_void_p, f64p, C.c_int, C.c_int]
        L.hx_mask_owned.argtypes = [C.c_void_p, f64p, C.c_int]
        L.hx_permute_scale.argtypes = [C.c_void_p, f64p, f64p]
        L.hx_unpermute.argtypes = [C.c_void_p, f64p, f64p]
        L.hx_info.argtypes = [C.c_void_p, i64p, f64p]
        L.hx_perm.argtypes = [C.c_void_p, i64p, i64p]
        L.hx_factor.restype = C.c_longlong
        L.hx_factor.argtypes = [C.c_void_p, f64p, C.c_void_p]
        L.hx_solve.argtypes = [C.c_void_p, f64p, f64p]
        L.hx_lsolve.argtypes = [C.c_void_p, f64p]
        L.hx_rsolve.argtypes = [C.c_void_p, f64p]
        L.hx_get_factors.argtypes = [C.c_void_p, i64p, i64p, f64p, i64p, i64p, f64p]
        L.hx_free.argtypes = [C.c_void_p]
        self.L = L

    def run(self, A, ordering=0, grid=None, leaf=0, relax=1, maxw=0, p=None, q=None, Rs=None, factor=True):
        import scipy.sparse as sp
        A = sp.csc_matrix(A); A.sort_indices()
        n = A.shape[0]
        Ap = A.indptr.astype(np.int64); Ai = A.indices.astype(np.int64); Ax = A.data.astype(np.float64)
        g = None if grid is None else np.array(list(grid) + [1] * (3 - len(grid)), dtype=np.int32)
        pp = None if p is None else np.ascontiguousarray(p, np.int64)
        qq = None if q is None else np.ascontiguousarray(q, np.int64)
        h = self.L.hx_create(n, Ap, Ai, ordering, None if g is None else g.ctypes.data_as(C.c_void_p), leaf,
                             relax, maxw, None if pp is None else pp.ctypes.data_as(C.c_void_p),
                             None if qq is None else qq.ctypes.data_as(C.c_void_p))
        assert h, "hx_create failed"
        try:
            info = np.zeros(16, np.int64); fl = np.zeros(2)
            self.L.hx_info(h, info, fl)
            keys = ("n", "nsn", "nlevels", "lu_size", "cb_size", "nnzL_exact", "nnzL_stored", "sum_r",
                    "max_front", "max_k", "max_children", "nranks", "lu_top_size", "cb_iface_size", "vbuf_len", "ntop")
            out = dict(zip(keys, (int(v) for v in info)))
            out["flops_exact"], out["flops_stored"] = float(fl[0]), float(fl[1])
            po = np.zeros(n, np.int64); qo = np.zeros(n, np.int64)
            self.L.hx_perm(h, po, qo)
            out["p"], out["q"] = po, qo
            if factor:
                rs = None if Rs is None else np.ascontiguousarray(Rs, np.float64)
                out["bad"] = int(self.L.hx_factor(h, Ax, None if rs is None else rs.ctypes.data_as(C.c_void_p)))
                nl = out["nnzL_exact"]
                Lp = np.zeros(n + 1, np.int64); Li = np.zeros(nl, np.int64); Lx = np.zeros(nl)
                Up = np.zeros(n + 1, np.int64); Ui = np.zeros(nl, np.int64); Ux = np.zeros(nl)
                self.L.hx_get_factors(h, Lp, Li, Lx, Up, Ui, Ux)
                out.update(Lp=Lp, Li=Li, Lx=Lx, Up=Up, Ui=Ui, Ux=Ux)

                def solve(b):
                    x = np.zeros(n)
                    self.L.hx_solve(h, np.ascontiguousarray(b, np.float64), x)
                    return x
                b = np.cos(np.arange(n) * 0.37) + 1.5
                out["b"], out["x"] = b, solve(b)
                y = b.copy(); self.L.hx_lsolve(h, y); out["lsolve_b"] = y
                z = b.copy(); self.L.hx_rsolve(h, z); out["rsolve_b"] = z
            return out
        finally:
            self.L.hx_free(h)


class PartitionedWalk:
    """One rank's view of the partitioned factorization / solve, walked on the CPU (tests/hostexec.cpp).
    `allreduce(array)` must sum a float64 numpy array in place over the ranks (gloo in the tests)."""

    def __init__(self, hx, A, nranks, rank, allreduce, **kw):
        import scipy.sparse as sp
        A = sp.csc_matrix(A); A.sort_indices()
        self.L, self.rank, self.nranks, self.allreduce = hx.L, rank, nranks, allreduce
        self.n = n = A.shape[0]
        self.Ap = A.indptr.astype(np.int64); self.Ai = A.indices.astype(np.int64)
        self.h = self.L.hx_create2(n, self.Ap, self.Ai, kw.get("ordering", 0), None, kw.get("leaf", 0), 1,
                                   kw.get("maxw", 0), None, None, nranks)
        assert self.h
        info = np.zeros(16, np.int64); fl = np.zeros(2)
        self.L.hx_info(self.h, info, fl)
        self.info = dict(zip(("n", "nsn", "nlevels", "lu_size", "cb_size", "nnzL_exact", "nnzL_stored", "sum_r",
                              "max_front", "max_k", "max_children", "nranks", "lu_top_size", "cb_iface_size",
                              "vbuf_len", "ntop"), (int(v) for v in info)))
        self.owner = np.zeros(self.info["nsn"], np.int64)
        self.L.hx_owner(self.h, self.owner)
        self.p = np.zeros(n, np.int64); self.q = np.zeros(n, np.int64)
        self.L.hx_perm(self.h, self.p, self.q)

    def _view(self, which, count):
        ln = C.c_longlong()
        ptr = self.L.hx_buffer(self.h, which, C.byref(ln))
        assert count <= ln.value
        if count == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array((C.c_double * count).from_address(ptr))

    def factor(self, Ax, Rs):
        Ax = np.ascontiguousarray(Ax, np.float64); Rs = np.ascontiguousarray(Rs, np.float64)
        self.L.hx_factor_phase(self.h, Ax, Rs.ctypes.data_as(C.c_void_p), self.rank, 0)
        if self.nranks == 1:
            return 0
        # exchange 1: the subtree roots' contribution blocks reach the column owners (the GPU path stores every column
        # straight into its owner's pool; here every rank's slots are zero except the producer's, so a sum delivers them)
        self.allreduce(self._view(1, self.info["cb_iface_size"]))
        lv = np.zeros(self.info["nlevels"], np.int64)
        nl = self.L.hx_top_levels(self.h, lv)
        segs = np.zeros(2 * self.info["nsn"] + 2, np.int64)
        lu = self._view(0, self.info["lu_size"])
        bad = -1
        for l in lv[:nl]:
            self.L.hx_top_panels(self.h, self.rank, int(l))
            # exchange 2 (per level): the panel owners publish P_s (non-owners hold zeros there)
            for i in range(self.L.hx_top_segments(self.h, int(l), 0, segs)):
                self.allreduce(lu[segs[2 * i]:segs[2 * i] + segs[2 * i + 1]])
            bad = int(self.L.hx_top_update(self.h, self.rank, int(l)))
        # exchange 3: every rank publishes the rows of U12' it owns (zeros elsewhere) -- the solves read all of T_s
        for l in lv[:nl]:
            for i in range(self.L.hx_top_segments(self.h, int(l), 1, segs)):
                self.allreduce(lu[segs[2 * i]:segs[2 * i] + segs[2 * i + 1]])
        return bad

    def solve(self, b):
        n = self.n
        w = np.zeros(n); x = np.zeros(n)
        self.L.hx_permute_scale(self.h, np.ascontiguousarray(b, np.float64), w)
        self.L.hx_lsolve_phase(self.h, w, self.rank, 0)
        if self.nranks > 1:
            self.allreduce(self._view(2, self.info["vbuf_len"]))
            self.L.hx_lsolve_phase(self.h, w, self.rank, 1)
            self.L.hx_rsolve_phase(self.h, w, self.rank, 1)
        self.L.hx_rsolve_phase(self.h, w, self.rank, 0)
        if self.nranks > 1:
            self.L.hx_mask_owned(self.h, w, self.rank)
            self.allreduce(w)
        self.L.hx_unpermute(self.h, w, x)
        return x

    def close(self):
        if self.h:
            self.L.hx_free(self.h); self.h = None


@pytest.fixture(scope="session")
def hostexec():
    return HostExec()


def relerr(a, b):
    """max |a-b| / max(|b|, tiny) elementwise -- the 'relative' in 'L/U entries within 1e-12 relative'."""
    a = np.asarray(a); b = np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relerr_csc(data, ref, indptr, floor=1e-2):
    """Entry-wise relative error of CSC values, with |ref_ij| floored at `floor` * (largest |entry| of
    its column): entries that are the residue of cancellation are compared against the column scale."""
    data = np.asarray(data); ref = np.asarray(ref)
    if data.size == 0:
        return 0.0
    counts = np.diff(indptr)
    col = np.repeat(np.arange(counts.size), counts)
    cmax = np.zeros(counts.size)
    np.maximum.at(cmax, col, np.abs(ref))
    denom = np.maximum(np.abs(ref), floor * cmax[col])
    return float(np.max(np.abs(data - ref) / np.maximum(denom, 1e-300)))
