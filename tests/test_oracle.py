"""CPU: pin the oracle as far as this environment allows.

The reference holds no golden vectors and cannot run here (no Julia, no UMFPACK): PARITY UNPINNED
(oracle/ref_lu.c header).  What can be checked: the contract the reference documents
(L*U == (Rs .* A)[p,q], src:307), the identities its tests assert (test/runtests.jl:51-186) at its
own tolerances, and agreement with an independent LU (SciPy SuperLU) under the same pivot order.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

TOL = 1.0e-12        # test/runtests.jl:25
DENSE_TOL = 1.0e-10  # test/runtests.jl:26


def isapprox(x, y, tol):
    # Julia isapprox(x, y; rtol, atol): norm(x-y) <= max(atol, rtol*max(norm(x), norm(y)))
    return np.linalg.norm(x - y) <= max(tol, tol * max(np.linalg.norm(x), np.linalg.norm(y)))


def contract_error(F, A):
    B = (sp.diags(F.Rs) @ A).tocsr()[F.p][:, F.q]
    return abs(F.L @ F.U - B).max()


@pytest.mark.parametrize("nel", [1, 2, 3, 5, 8, 13, 40, 200])
def test_fe_fixture_contract_and_solves(O, W, nel):
    A = W.fe_test_matrix(nel, seed=100 + nel)
    n = A.shape[0]
    F = O.OracleLU(A, Rs=O.row_scale_sum(A))
    assert F.bad_col == -1
    assert contract_error(F, A) < 1e-12
    assert np.array_equal(np.sort(F.p), np.arange(n)) and np.array_equal(np.sort(F.q), np.arange(n))
    assert np.allclose(F.L.diagonal(), 1.0) and abs(sp.triu(F.L, 1)).sum() == 0 and abs(sp.tril(F.U, -1)).sum() == 0
    b = W.rhs(n, 7)
    # lsolve!/rsolve! identities (test:51,70,86,104) and ldiv! vs an independent solve (test:163)
    assert isapprox(F.L @ F.lsolve(b), b, TOL)
    assert isapprox(F.U @ F.usolve(b), b, DENSE_TOL)
    assert isapprox(F.solve(b), spla.spsolve(A, b), 1e-10 if nel > 50 else TOL * 100)
    # the reference's dense-chunk algorithm restated (src:101-243, 349-392) gives the same answers
    RC = O.RefChunks(F.L, F.U)
    assert RC.total_chunks == (n + min(8, n) - 1) // min(8, n)
    assert isapprox(RC.lsolve(b), F.lsolve(b), TOL)
    assert isapprox(RC.rsolve(b), F.usolve(b), DENSE_TOL)
    assert isapprox(RC.ldiv(F.p, F.q, F.Rs, b), F.solve(b), TOL)
    RC.close()


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 16, 17, 64, 200])
def test_dense_fixture(O, W, n):
    A = W.dense_random(n, seed=n)
    F = O.OracleLU(A, Rs=O.row_scale_sum(A))
    assert F.bad_col == -1
    assert contract_error(F, A) < 1e-13
    b = W.rhs(n, 11)
    x = F.solve(b)
    assert isapprox(x, np.linalg.solve(A.toarray(), b), DENSE_TOL)
    RC = O.RefChunks(F.L, F.U)
    assert isapprox(RC.ldiv(F.p, F.q, F.Rs, b), x, DENSE_TOL)
    # chunk ranges (src:111-143), 0-based half-open
    rg = RC.ranges()
    cs = min(8, n)
    T = RC.total_chunks
    assert np.array_equal(rg["lc0"], np.arange(T) * cs)
    assert np.array_equal(rg["uc0"], (T - 1 - np.arange(T)) * cs)
    assert np.all(rg["lr1"] == n) or n <= cs       # dense: rectangular part reaches the last row
    assert np.all(rg["ur0"] == 0) or n <= cs
    RC.close()


def test_chunks_dimension_mismatch(O, W):
    A = W.fe_test_matrix(3)
    F = O.OracleLU(A)
    RC = O.RefChunks(F.L, F.U)
    with pytest.raises(ValueError):                 # src:288-290
        RC.ldiv(F.p, F.q, F.Rs, np.ones(A.shape[0] + 1))
    RC.close()


def test_chunk_size_clamped_and_custom(O, W):
    A = W.fe_test_matrix(1)                         # n = 5 < 8: chunk size clamps to n (src:72)
    F = O.OracleLU(A)
    assert O.RefChunks(F.L, F.U).total_chunks == 1
    A = W.fe_test_matrix(10)
    F = O.OracleLU(A)
    b = W.rhs(A.shape[0], 3)
    for cs in (1, 3, 8, 41, 100):
        RC = O.RefChunks(F.L, F.U, cs)
        assert isapprox(RC.ldiv(F.p, F.q, F.Rs, b), F.solve(b), TOL)
        RC.close()


@pytest.mark.parametrize("shape", [(12, 12), (31, 17)])
def test_static_mode_matches_superlu(O, W, shape):
    """Same pivot order => same L, U as an independent implementation (SciPy SuperLU)."""
    A = W.laplacian_2d(*shape)
    n = A.shape[0]
    ident = np.arange(n)
    F = O.OracleLU(A, p=ident, q=ident)
    lu = spla.splu(A, permc_spec="NATURAL", diag_pivot_thresh=0, options=dict(SymmetricMode=True))
    assert np.array_equal(lu.perm_r, ident) and np.array_equal(lu.perm_c, ident)
    assert F.L.nnz == lu.L.nnz and F.U.nnz == lu.U.nnz
    assert abs(F.L - lu.L).max() < 1e-14 and abs(F.U - lu.U).max() < 1e-13


def test_partial_pivot_matches_superlu_perms(O, W):
    """With threshold 1.0 and no diagonal preference both codes do classical partial pivoting."""
    A = W.dense_random(30, seed=5)
    F = O.OracleLU(A, diag_tol=2.0)                 # >1 => diagonal never preferred
    lu = spla.splu(A, permc_spec="NATURAL", diag_pivot_thresh=1.0)
    # SciPy: L U = A[argsort(perm_r)][:, argsort(perm_c)]  
    assert np.array_equal(F.p, np.argsort(lu.perm_r))
    assert abs(F.L - lu.L).max() < 1e-12 and abs(F.U - lu.U).max() < 1e-12


def test_singular_is_reported(O):
    A = sp.csc_matrix(np.array([[1.0, 2.0], [2.0, 4.0]]))
    F = O.OracleLU(A, p=np.arange(2), q=np.arange(2))
    assert F.bad_col == 1


def test_row_scale_sum(O, W):
    A = W.fe_test_matrix(4)
    assert np.allclose(O.row_scale_sum(A), 1.0 / np.asarray(abs(A).sum(axis=1)).ravel())


def test_splitmix_known_answers(W):
    # splitmix64 reference outputs for seed 0: first state 0x9E3779B97F4A7C15 -> 0xE220A8397B1DCDAF
    v = W.splitmix64(0, 2)
    assert v[0] == (0xE220A8397B1DCDAF >> 11) / 2.0**53
    assert v[1] == (0x6E789E6AA1B965F4 >> 11) / 2.0**53


# ---------------------------------------------------------------------------------------------------------------
# The threaded CPU multifrontal port (oracle/ref_mf.cpp) that bench.py times as cpu_baseline / --impl reference:
# pinned against the left-looking oracle entry by entry.
@pytest.mark.parametrize("case", ["lap2d_30", "lap3d_14", "block_border", "lap3d_24_blas"])
def test_cpu_multifrontal_port_matches_oracle(O, W, case):
    A = {"lap2d_30": lambda: W.laplacian_2d(30), "lap3d_14": lambda: W.laplacian_3d(14),
         "block_border": lambda: W.block_border(nblocks=4, nel=5, ngr=5, border=8),
         "lap3d_24_blas": lambda: W.laplacian_3d(24)}[case]()
    n = A.shape[0]
    F = O.RefMF(A, threads=3)
    assert F.lu_(A.data) == -1
    p, q = F.perm()
    L, U, Rs = F.factors()
    assert np.array_equal(Rs, O.row_scale_sum(A))
    ref = O.OracleLU(A, p=p, q=q, Rs=Rs)
    assert np.array_equal(L.indices, ref.Li) and np.array_equal(U.indices, ref.Ui)
    assert abs(L - ref.L).max() < 1e-13 and abs(U - ref.U).max() < 1e-13 * abs(ref.U).max()
    b = W.rhs(n, 47)
    x = F.ldiv(b)
    assert np.linalg.norm(x - ref.solve(b)) <= 1e-12 * np.linalg.norm(x)
    vals = A.data * 1.01                                  # refactorization with the analysis reused
    assert F.lu_(vals) == -1
    A2 = A.copy(); A2.data = vals
    x2 = F.ldiv(b)
    assert np.linalg.norm(A2 @ x2 - b) <= 1e-13 * np.linalg.norm(b)
    assert F.growth <= 1.0 + 1e-12                         # diagonally dominant: no multiplier above 1
    F.close()


def test_bench_poisson_checker_is_exact(W):
    """bench.py's independent full-size checker (DST-I diagonalisation) against a sparse direct solve."""
    import importlib.util, os, scipy.sparse as sp, scipy.sparse.linalg as spla
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    for kind, e in (("lap3d", 11), ("lap2d", 37)):
        A = bench.make_matrix(W, kind, e)
        n = A.shape[0]
        b = W.rhs(n, 5)
        x = bench.poisson_solve(kind, e, 2e-3, b)
        xs = spla.spsolve(sp.csc_matrix(A + 2e-3 * sp.identity(n)), b)
        assert np.linalg.norm(x - xs) <= 1e-13 * np.linalg.norm(xs)
