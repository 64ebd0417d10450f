"""Known-answer fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py).

CPU: the oracle and the host symbolic layout (walked on the CPU by tests/hostexec.cpp) reproduce the
committed L, U, x.  GPU: the CUDA path, given the fixture's (p, q, Rs), reproduces them through the
C ABI.  The fixtures come from the oracle cross-checked with SciPy SuperLU -- the reference itself
holds no golden vectors and cannot run here (parity unpinned)."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, relerr_csc

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
NAMES = [os.path.basename(f)[:-4] for f in FILES]


def load(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    n = g["Ap"].size - 1
    A = sp.csc_matrix((g["Ax"], g["Ai"], g["Ap"]), shape=(n, n))
    return g, A, n


def test_fixtures_present():
    assert len(FILES) >= 5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(O, name):
    g, A, n = load(name)
    F = O.OracleLU(A, p=g["p"], q=g["q"], Rs=g["Rs"])
    assert np.array_equal(O.row_scale_sum(A), g["Rs"])
    assert np.array_equal(F.Lp, g["Lp"]) and np.array_equal(F.Li, g["Li"])
    assert np.array_equal(F.Up, g["Up"]) and np.array_equal(F.Ui, g["Ui"])
    assert relerr_csc(F.Lx, g["Lx"], g["Lp"]) < 1e-14 and relerr_csc(F.Ux, g["Ux"], g["Up"]) < 1e-14
    assert np.linalg.norm(F.solve(g["b"]) - g["x"]) <= 1e-14 * np.linalg.norm(g["x"])
    RC = O.RefChunks(F.L, F.U)                 # the reference's dense-chunk solve (src:286-392)
    assert np.linalg.norm(RC.ldiv(F.p, F.q, F.Rs, g["b"]) - g["chunks_x"]) <= 1e-14 * np.linalg.norm(g["x"])
    RC.close()


@pytest.mark.parametrize("name", NAMES)
def test_host_layout_reproduces_golden(hostexec, O, name):
    g, A, n = load(name)
    out = hostexec.run(A, ordering=2, p=g["p"], q=g["q"], Rs=g["Rs"])
    assert out["bad"] == -1
    # a given (p, q) comes back unchanged: the analysis may relabel internally by a postorder of the elimination
    # tree, but everything at the boundary is reported in the caller's labelling
    assert np.array_equal(out["p"], g["p"]) and np.array_equal(out["q"], g["q"])
    L = sp.csc_matrix((out["Lx"], out["Li"], out["Lp"]), shape=(n, n))
    U = sp.csc_matrix((out["Ux"], out["Ui"], out["Up"]), shape=(n, n))
    B = (sp.diags(g["Rs"]) @ A).tocsr()[out["p"]][:, out["q"]]
    assert abs(L @ U - B).max() < 1e-13 * max(1.0, abs(B).max())
    assert np.array_equal(out["Lp"], g["Lp"]) and np.array_equal(out["Li"], g["Li"])
    assert np.array_equal(out["Up"], g["Up"]) and np.array_equal(out["Ui"], g["Ui"])
    assert relerr_csc(out["Lx"], g["Lx"], g["Lp"]) < 1e-12 and relerr_csc(out["Ux"], g["Ux"], g["Up"]) < 1e-12
    assert np.linalg.norm(out["lsolve_b"] - O.csc_lsolve(sp.csc_matrix((g["Lx"], g["Li"], g["Lp"]), shape=(n, n)), out["b"])) <= 1e-12 * np.linalg.norm(out["lsolve_b"])
    assert np.linalg.norm(out["rsolve_b"] - O.csc_usolve(sp.csc_matrix((g["Ux"], g["Ui"], g["Up"]), shape=(n, n)), out["b"])) <= 1e-10 * np.linalg.norm(out["rsolve_b"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_golden(smslu, name):
    g, A, n = load(name)
    F = smslu.ParallelSparseLU(A, p=g["p"], q=g["q"], Rs=g["Rs"])
    x = np.empty(n)
    smslu.ldiv_(x, F, g["b"])
    assert np.linalg.norm(x - g["x"]) <= 1e-12 * np.linalg.norm(g["x"])
    L, U = F.L, F.U
    B = (sp.diags(g["Rs"]) @ A).tocsr()[F.p][:, F.q]
    assert abs(L @ U - B).max() < 1e-13 * max(1.0, abs(B).max())
    # bit-exact permutations and structure, entries to 1e-12, unconditionally (the given p, q come back unchanged)
    assert np.array_equal(F.p, g["p"]) and np.array_equal(F.q, g["q"])
    assert np.array_equal(L.indptr, g["Lp"]) and np.array_equal(L.indices, g["Li"])
    assert np.array_equal(U.indptr, g["Up"]) and np.array_equal(U.indices, g["Ui"])
    assert relerr_csc(L.data, g["Lx"], g["Lp"]) < 1e-12 and relerr_csc(U.data, g["Ux"], g["Up"]) < 1e-12
    y = g["b"].copy(); smslu.lsolve_(F, y)
    assert np.linalg.norm(y - g["lsolve_b"]) <= 1e-12 * np.linalg.norm(g["lsolve_b"])
    y = g["b"].copy(); smslu.rsolve_(F, y)
    assert np.linalg.norm(y - g["usolve_b"]) <= 1e-10 * np.linalg.norm(g["usolve_b"])
    F.close()
