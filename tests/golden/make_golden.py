#!/usr/bin/env python
"""Generates tests/golden/*.npz -- small known-answer fixtures for the hot path.

The reference holds no golden vectors and cannot run here (Julia + UMFPACK are absent), so these
fixtures are produced by the CPU oracle (oracle/ref_lu.c, oracle/ref_chunks.c) and only kept when an
independent factorization (SciPy SuperLU forced to the same pivot order) reproduces L and U entry
by entry.  They pin the oracle, the host symbolic analysis and the CUDA path against drift; they do
NOT pin parity with UMFPACK (parity unpinned).

    python tests/golden/make_golden.py          # rewrites the .npz files next to this script
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import smslu  # noqa: E402  (host-only use: workloads + symbolic analysis)
from sharedmemsparselu_jl_b200 import _SymbolicOnly, workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402


def superlu_same_pivots(A, p, q, Rs):
    """SciPy SuperLU on the pre-permuted, pre-scaled matrix with natural ordering and diagonal pivots."""
    B = sp.csc_matrix((sp.diags(Rs) @ A).tocsr()[p][:, q])
    lu = spla.splu(B, permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    assert np.array_equal(lu.perm_r, np.arange(A.shape[0])) and np.array_equal(lu.perm_c, np.arange(A.shape[0]))
    return sp.csc_matrix(lu.L), sp.csc_matrix(lu.U)


CASES = {
    # name: (matrix, how the pivot order is fixed)
    "lap2d_12x9_nd": (lambda: W.laplacian_2d(12, 9), "native"),
    "lap3d_6_nd": (lambda: W.laplacian_3d(6), "native"),
    "fe_nel6_pivoted": (lambda: W.fe_test_matrix(6, seed=11), "oracle_pivoting"),       # reference fixture, test:12-21
    "dense_12_pivoted": (lambda: W.dense_random(12, seed=12), "oracle_partial"),         # rand(n,n), test:41-42
    "block_border_3x2": (lambda: W.block_border(nblocks=3, nel=2, ngr=5, border=4), "native"),
}


def main():
    for name, (mk, how) in CASES.items():
        A = mk()
        n = A.shape[0]
        Rs = O.row_scale_sum(A)
        if how == "native":
            S = _SymbolicOnly(A)
            p, q = S.p.copy(), S.q.copy()
            S.close()
        elif how == "oracle_pivoting":
            piv = O.OracleLU(A, Rs=Rs)
            p, q = piv.p, piv.q
        else:
            piv = O.OracleLU(A, Rs=Rs, diag_tol=2.0)
            p, q = piv.p, piv.q
        F = O.OracleLU(A, p=p, q=q, Rs=Rs)
        assert F.bad_col == -1
        L2, U2 = superlu_same_pivots(A, p, q, Rs)
        L2.sort_indices(); U2.sort_indices()
        Ld, Ud = F.L.toarray(), F.U.toarray()
        errL = np.max(np.abs(Ld - L2.toarray())); errU = np.max(np.abs(Ud - U2.toarray()))
        assert errL < 1e-13 and errU < 1e-13, (name, errL, errU)
        b = W.rhs(n, 47)
        x = F.solve(b)
        RC = O.RefChunks(F.L, F.U)
        xc = RC.ldiv(F.p, F.q, F.Rs, b)        # the reference's own dense-chunk solve on the same factors
        assert np.linalg.norm(xc - x) <= 1e-13 * np.linalg.norm(x)
        assert np.linalg.norm(A @ x - b) <= 1e-12 * np.linalg.norm(b)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            Ap=A.indptr.astype(np.int64), Ai=A.indices.astype(np.int64), Ax=A.data,
            p=p.astype(np.int64), q=q.astype(np.int64), Rs=Rs,
            Lp=F.Lp, Li=F.Li, Lx=F.Lx, Up=F.Up, Ui=F.Ui, Ux=F.Ux,
            b=b, x=x, lsolve_b=F.lsolve(b), usolve_b=F.usolve(b), chunks_x=xc)
        print("%-20s n=%4d nnzL=%5d  |L-L_superlu|=%.1e |U-U_superlu|=%.1e" % (name, n, F.Li.size, errL, errU))


if __name__ == "__main__":
    main()
