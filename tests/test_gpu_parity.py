"""GPU: parity of the CUDA path (through the C ABI) with the CPU oracle, on the same inputs.

Bar (BASELINE.json north_star): bit-exact symbolic structure, permutations and pivots; L/U entries
within 1e-12 relative; solve residual no worse than the oracle's; lsolve!/rsolve!/ldiv! at the
reference's own tolerances (test/runtests.jl:25-26: 1e-12, 1e-10 for dense / rsolve)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import relerr, relerr_csc

pytestmark = pytest.mark.gpu

TOL = 1.0e-12
DENSE_TOL = 1.0e-10


def isapprox(x, y, tol):
    return np.linalg.norm(x - y) <= max(tol, tol * max(np.linalg.norm(x), np.linalg.norm(y)))


def residual(A, x, b):
    return np.linalg.norm(A @ x - b) / np.linalg.norm(b)


CASES = {
    "n1": lambda W: (sp.csc_matrix(np.array([[2.5]])), {}),
    "diag7": lambda W: (sp.identity(7, format="csc") * 3.0, {}),
    "lap2d_10": lambda W: (W.laplacian_2d(10), {}),
    "lap2d_37x23": lambda W: (W.laplacian_2d(37, 23), {}),
    "lap2d_37x23_grid": lambda W: (W.laplacian_2d(37, 23), dict(ordering="nd_grid", grid=(37, 23))),
    "lap2d_64_norelax": lambda W: (W.laplacian_2d(64), dict(relax=False)),
    "lap2d_64_w8": lambda W: (W.laplacian_2d(64), dict(max_width=8)),
    "lap2d_100_config1": lambda W: (W.laplacian_2d(100), {}),          # BASELINE configs[0]
    "lap2d_200": lambda W: (W.laplacian_2d(200), {}),
    "lap3d_12": lambda W: (W.laplacian_3d(12), {}),
    "lap3d_20": lambda W: (W.laplacian_3d(20), {}),
    "natural_15": lambda W: (W.laplacian_2d(15), dict(ordering="natural")),
    "fe_50_shifted": lambda W: (sp.csc_matrix(W.fe_test_matrix(50, seed=3) + 5 * sp.identity(201)), {}),
    "dense_40_dd": lambda W: (sp.csc_matrix(W.dense_random(40, seed=2) + 40 * sp.identity(40)), {}),
    "dense_150_dd": lambda W: (sp.csc_matrix(W.dense_random(150, seed=4) + 150 * sp.identity(150)), {}),
    "block_border_small": lambda W: (W.block_border(nblocks=4, nel=5, ngr=5, border=8), {}),
}


def unsym_random(W, n, per_row, seed):
    """Structurally unsymmetric random pattern (the symbolic analysis works on A+A'), strictly diagonally dominant."""
    u = W.splitmix64(seed, 2 * n * per_row)
    rows = np.repeat(np.arange(n), per_row)
    cols = np.minimum((u[:n * per_row] * n).astype(np.int64), n - 1)
    vals = u[n * per_row:] - 0.5
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A = A - sp.diags(A.diagonal())
    A = A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
    A = sp.csc_matrix(A); A.sort_indices()
    A.indptr = A.indptr.astype(np.int64); A.indices = A.indices.astype(np.int64)
    return A


@pytest.mark.parametrize("name", list(CASES))
def test_factor_and_solve_match_oracle(smslu, O, W, name):
    A, kw = CASES[name](W)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A, **kw)
    p, q, Rs = F.p, F.q, F.Rs
    # permutations valid; row scaling bit-identical to the oracle's (same summation order)
    assert np.array_equal(np.sort(p), np.arange(n)) and np.array_equal(np.sort(q), np.arange(n))
    assert np.array_equal(Rs, O.row_scale_sum(A))
    ref = O.OracleLU(A, p=p, q=q, Rs=Rs)
    assert ref.bad_col == -1
    L, U = F.L, F.U
    # bit-exact structure
    assert np.array_equal(L.indptr, ref.Lp) and np.array_equal(L.indices, ref.Li)
    assert np.array_equal(U.indptr, ref.Up) and np.array_equal(U.indices, ref.Ui)
    # entries within 1e-12 relative
    assert relerr_csc(L.data, ref.Lx, ref.Lp) < 1e-12
    assert relerr_csc(U.data, ref.Ux, ref.Up) < 1e-12
    if name.startswith("lap"):          # M-matrices: no cancellation, plain entry-wise relative error
        assert relerr(L.data, ref.Lx) < 1e-12 and relerr(U.data, ref.Ux) < 1e-12
    # contract L*U == (Rs .* A)[p,q]  (src:307)
    B = (sp.diags(Rs) @ A).tocsr()[p][:, q]
    assert abs(L @ U - B).max() < 1e-13 * max(1.0, abs(B).max())
    # lsolve! / rsolve! / ldiv!  (test:51,70,86,104,163)
    b = W.rhs(n, 47)
    x = b.copy(); smslu.lsolve_(F, x)
    assert isapprox(x, ref.lsolve(b), TOL)
    x = b.copy(); smslu.rsolve_(F, x)
    assert isapprox(x, ref.usolve(b), DENSE_TOL)
    x = np.empty(n); b0 = b.copy()
    assert smslu.ldiv_(x, F, b) is x
    assert np.array_equal(b, b0)                                   # b untouched (src:320-321)
    xo = ref.solve(b)
    assert isapprox(x, xo, TOL if n < 5000 else 1e-11)
    assert residual(A, x, b) <= max(4 * residual(A, xo, b), 1e-15)
    # new right-hand side (test:166-169)
    b2 = W.rhs(n, 48)
    smslu.ldiv_(x, F, b2)
    assert isapprox(x, ref.solve(b2), TOL if n < 5000 else 1e-11)
    smslu.cleanup_ParallelSparseLU_(F)


@pytest.mark.parametrize("n,per_row,kw", [(400, 6, {}), (1500, 9, dict(max_width=64)), (3000, 4, dict(relax=False))])
def test_structurally_unsymmetric_pattern(smslu, O, W, n, per_row, kw):
    """The symbolic analysis works on pattern(A + A'), so for a structurally unsymmetric A the stored factors are a
    superset of the oracle's exact (Gilbert-Peierls) structure, padded with explicit zeros: same values where the oracle
    has entries, zero elsewhere, same contract L*U == (Rs .* A)[p,q] (src:307), same solves."""
    A = unsym_random(W, n, per_row, seed=7 + n)
    assert abs(abs(A) > 0).astype(np.int8).T.tocsc().nnz == A.nnz and ((abs(A) > 0) != (abs(A.T) > 0)).nnz > 0
    F = smslu.ParallelSparseLU(A, **kw)
    ref = O.OracleLU(A, p=F.p, q=F.q, Rs=F.Rs)
    L, U = F.L, F.U
    B = (sp.diags(F.Rs) @ A).tocsr()[F.p][:, F.q]
    assert abs(L @ U - B).max() < 1e-13 * max(1.0, abs(B).max())
    assert abs(L - ref.L).max() < 1e-12 and abs(U - ref.U).max() < 1e-12 * abs(ref.U).max()
    assert L.nnz >= ref.L.nnz and U.nnz >= ref.U.nnz
    b = W.rhs(n, 3)
    x = np.empty(n); smslu.ldiv_(x, F, b)
    assert isapprox(x, ref.solve(b), TOL)
    y = b.copy(); smslu.lsolve_(F, y)
    assert isapprox(y, ref.lsolve(b), TOL)
    y = b.copy(); smslu.rsolve_(F, y)
    assert isapprox(y, ref.usolve(b), DENSE_TOL)
    F.close()


@pytest.mark.parametrize("name", ["lap2d_37x23", "lap3d_12", "fe_50_shifted", "lap2d_200"])
@pytest.mark.parametrize("mode", ["shift", "random"])
def test_refactor_with_new_values(smslu, O, W, name, mode):
    """lu!(F, A) with the same pattern and new values (src:245-279, test:171-186), repeatedly.
    'shift' is BASELINE config 2's rule (A*(1+0.01k) + k*1e-3*I: stays an M-matrix, no cancellation,
    so entries agree to 1e-12); 'random' perturbs every entry by up to 15%, which makes entries
    cancel -- there two correct eliminations with different summation orders (left-looking oracle vs
    multifrontal, also on the CPU: tests/hostexec.cpp) differ by cond(A)*eps, so the bar is the
    looser 1e-8 on entries and the reference's own solve tolerance on x."""
    A, kw = CASES[name](W)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A, **kw)
    p, q = F.p.copy(), F.q.copy()
    for k in range(1, 4):
        A2 = A.copy()
        if mode == "shift":
            A2.data = A.data * (1.0 + 0.01 * k)
        else:
            A2.data = A.data * (1.0 + 0.05 * k * W.splitmix64(900 + k, A.nnz))
        A2 = sp.csc_matrix(A2 + k * 1e-3 * sp.identity(n))
        assert smslu.lu_(F, A2) is None
        assert np.array_equal(F.p, p) and np.array_equal(F.q, q)     # static pivot order
        ref = O.OracleLU(A2, p=p, q=q, Rs=F.Rs)
        assert np.array_equal(F.L.indices, ref.Li)
        tol = 1e-12 if mode == "shift" and not name.startswith("fe") else 1e-8
        assert relerr_csc(F.L.data, ref.Lx, ref.Lp) < tol and relerr_csc(F.U.data, ref.Ux, ref.Up) < tol
        b = W.rhs(n, 60 + k)
        x = np.empty(n)
        smslu.ldiv_(x, F, b)
        xo = ref.solve(b)
        assert isapprox(x, xo, TOL if mode == "shift" and n < 5000 else 1e-9)
        assert residual(A2, x, b) <= max(4 * residual(A2, xo, b), 1e-15)
    F.close()


@pytest.mark.parametrize("nel", [1, 2, 3, 7, 20, 64, 200])
def test_reference_sparse_testsets_with_given_pivots(smslu, O, W, nel):
    """The reference's 'lsolve!/rsolve!/sparse matrix' testsets (test:55-106, 148-188) on its
    test_matrix fixture, with (p, q, Rs) handed in as the Julia shim hands in UMFPACK's."""
    A = W.fe_test_matrix(nel, seed=nel)
    n = A.shape[0]
    piv = O.OracleLU(A, Rs=O.row_scale_sum(A))      # stand-in for UMFPACK's pivot search
    F = smslu.ParallelSparseLU(A, p=piv.p, q=piv.q, Rs=piv.Rs)
    ref = O.OracleLU(A, p=F.p, q=F.q, Rs=piv.Rs)
    b = W.rhs(n, 5)
    # as in the reference: lsolve!(F, x) against `F.L \ b` with F's own factor (test:68-70, 102-104)
    x = b.copy(); smslu.lsolve_(F, x)
    assert isapprox(x, O.csc_lsolve(F.L, b), TOL)
    x = b.copy(); smslu.rsolve_(F, x)
    assert isapprox(x, O.csc_usolve(F.U, b), DENSE_TOL)
    x = np.empty(n); smslu.ldiv_(x, F, b)
    assert isapprox(x, ref.solve(b), TOL * 10)
    assert isapprox(x, np.linalg.solve(A.toarray(), b), 1e-9)
    F.close()


@pytest.mark.parametrize("n", [1, 2, 5, 8, 9, 33, 100, 200])
def test_reference_dense_testsets_with_given_pivots(smslu, O, W, n):
    """'lsolve!/rsolve!/dense matrix' testsets (test:38-53, 74-88, 108-146) on rand(n,n)."""
    A = W.dense_random(n, seed=1000 + n)
    piv = O.OracleLU(A, Rs=O.row_scale_sum(A), diag_tol=2.0)     # classical partial pivoting
    F = smslu.ParallelSparseLU(A, p=piv.p, q=piv.q, Rs=piv.Rs)
    ref = O.OracleLU(A, p=F.p, q=F.q, Rs=piv.Rs)
    assert relerr_csc(F.L.data, ref.Lx, ref.Lp) < 1e-10 and relerr_csc(F.U.data, ref.Ux, ref.Up) < 1e-10
    b = W.rhs(n, 6)
    x = b.copy(); smslu.lsolve_(F, x)
    assert isapprox(x, O.csc_lsolve(F.L, b), TOL)               # test:49-51
    x = b.copy(); smslu.rsolve_(F, x)
    assert isapprox(x, O.csc_usolve(F.U, b), DENSE_TOL)         # test:84-86
    x = np.empty(n); smslu.ldiv_(x, F, b)
    assert isapprox(x, np.linalg.solve(A.toarray(), b), DENSE_TOL * 10)
    F.close()


def test_errors(smslu, W):
    A = W.laplacian_2d(6)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A)
    with pytest.raises(smslu.DimensionMismatch):                    # src:289-290
        smslu.ldiv_(np.empty(n + 1), F, np.ones(n))
    with pytest.raises(smslu.DimensionMismatch):
        smslu.ldiv_(np.empty(n), F, np.ones(n - 1))
    with pytest.raises(smslu.SmsluError):                           # different pattern
        smslu.lu_(F, W.laplacian_2d(6, 6, 0.0)[:, ::-1].tocsc())
    S = sp.csc_matrix(np.array([[1.0, 2.0], [2.0, 4.0]]))
    with pytest.raises(smslu.SingularException):
        smslu.ParallelSparseLU(S, ordering="natural", scaling="none")
    F.close()


def test_device_resident_vectors(smslu, W):
    import torch
    A = W.laplacian_2d(40)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A)
    b = W.rhs(n, 3)
    xh = np.empty(n); smslu.ldiv_(xh, F, b)
    bd = torch.from_numpy(b).cuda(); xd = torch.empty(n, dtype=torch.float64, device="cuda")
    smslu.ldiv_(xd, F, bd)
    torch.cuda.synchronize()
    assert np.array_equal(xd.cpu().numpy(), xh)                     # deterministic, same bits
    vd = torch.from_numpy(np.ascontiguousarray(A.data)).cuda()
    smslu.lu_(F, vd)                                                # device-resident nzval
    smslu.ldiv_(xd, F, bd)
    torch.cuda.synchronize()
    assert np.array_equal(xd.cpu().numpy(), xh)
    xp = smslu.pinned_empty(n); bp = smslu.pinned_empty(n); bp[:] = b
    smslu.ldiv_(xp, F, bp)
    assert np.array_equal(np.asarray(xp), xh)
    F.close()


@pytest.mark.parametrize("case,nrhs", [("lap3d_10", 5), ("lap2d_150", 13), ("lap3d_14", 21), ("block_border_small", 8),
                                       ("lap3d_14", 37), ("lap2d_150", 70), ("block_border_small", 33), ("lap3d_20", 32)])
def test_multiple_rhs(smslu, O, W, case, nrhs):
    """Matrix right-hand sides (BASELINE config 5): the solve kernels sweep 8 / 4 / 1 columns at a time over
    interleaved work vectors, and from 9 columns on 32 at a time on the FP64 tensor pipe (k_fwd32 / k_bwd32; the
    remainder again 8 / 4 / 1); every column must equal the single-vector ldiv!/lsolve!/rsolve! of that column
    (1e-14 for the FMA sweeps, whose arithmetic is the single vector's; 1e-12, the reference's tolerance, for the
    tensor-pipe sweeps, which sum in a different order)."""
    A = {"lap3d_10": lambda: W.laplacian_3d(10), "lap2d_150": lambda: W.laplacian_2d(150),
         "lap3d_14": lambda: W.laplacian_3d(14), "lap3d_20": lambda: W.laplacian_3d(20),
         "block_border_small": lambda: W.block_border(nblocks=4, nel=5, ngr=5, border=8)}[case]()
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A)
    B = W.rhs(n, 47, nrhs=nrhs)
    X = np.empty((n, nrhs), order="F")
    B0 = B.copy(order="F")
    smslu.ldiv_(X, F, B)
    assert np.array_equal(B, B0)
    YL = B.copy(order="F"); smslu.lsolve_(F, YL)
    YU = B.copy(order="F"); smslu.rsolve_(F, YU)
    tol = 1e-14 if nrhs < 9 else 1e-12
    for r in range(nrhs):
        bcol = np.ascontiguousarray(B[:, r])
        x = np.empty(n); smslu.ldiv_(x, F, bcol)
        assert np.linalg.norm(X[:, r] - x) <= tol * np.linalg.norm(x), (r, np.linalg.norm(X[:, r] - x) / np.linalg.norm(x))
        assert residual(A, X[:, r], B[:, r]) < 1e-12
        y = bcol.copy(); smslu.lsolve_(F, y)
        assert np.linalg.norm(YL[:, r] - y) <= tol * np.linalg.norm(y), (r, "lsolve", np.linalg.norm(YL[:, r] - y) / np.linalg.norm(y))
        y = bcol.copy(); smslu.rsolve_(F, y)
        assert np.linalg.norm(YU[:, r] - y) <= tol * np.linalg.norm(y), (r, "rsolve", np.linalg.norm(YU[:, r] - y) / np.linalg.norm(y))
    F.close()


@pytest.mark.parametrize("n,nrhs", [(200, 33), (131, 12), (40, 70)])
def test_wide_sweeps_on_dense_matrices_with_given_pivots(smslu, O, W, n, nrhs):
    """The 32-wide tensor-pipe sweeps on the reference's dense fixture family (rand(n,n), partial pivoting handed in as
    GIVEN pivots => relabelled internally, chains of 128-column fronts with few rows): ldiv!/lsolve!/rsolve! of a block of
    right-hand sides against dense solves, at the reference's dense tolerance (test:25-26)."""
    A = W.dense_random(n, seed=2000 + n)
    piv = O.OracleLU(A, Rs=O.row_scale_sum(A), diag_tol=2.0)
    F = smslu.ParallelSparseLU(A, p=piv.p, q=piv.q, Rs=piv.Rs)
    B = W.rhs(n, 11, nrhs=nrhs)
    X = np.empty((n, nrhs), order="F")
    smslu.ldiv_(X, F, B)
    Xref = np.linalg.solve(A.toarray(), B)
    assert np.linalg.norm(X - Xref) <= DENSE_TOL * 10 * np.linalg.norm(Xref)
    L, U = F.L.toarray(), F.U.toarray()
    Y = B.copy(order="F"); smslu.lsolve_(F, Y)
    assert np.linalg.norm(Y - np.linalg.solve(L, B)) <= DENSE_TOL * np.linalg.norm(Y)
    Y = B.copy(order="F"); smslu.rsolve_(F, Y)
    Yref = np.linalg.solve(U, B)
    assert np.linalg.norm(Y - Yref) <= DENSE_TOL * 10 * np.linalg.norm(Yref)
    F.close()


def test_wide_sweeps_device_resident_blocks(smslu, W):
    """Device-resident blocks of right-hand sides (ld = n) through the wide sweeps, twice (the partial-sum scratch and
    arrival counters of the backward kernel are reused), bitwise equal run to run and equal to the host-buffer call."""
    import torch
    A = W.laplacian_3d(16)
    n, nrhs = A.shape[0], 45
    F = smslu.ParallelSparseLU(A)
    B = W.rhs(n, 5, nrhs=nrhs)
    Xh = np.empty((n, nrhs), order="F"); smslu.ldiv_(Xh, F, B)
    Bd = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()          # (nrhs, n) row-major == (n, nrhs) column-major
    Xd = torch.empty_like(Bd)
    from sharedmemsparselu_jl_b200 import _capi
    import ctypes as C
    outs = []
    for rep in range(2):
        _capi.check(F._h, _capi.lib().smslu_solve(F._h, C.c_void_p(Xd.data_ptr()), n, C.c_void_p(Bd.data_ptr()), n, nrhs, n, n))
        torch.cuda.synchronize()
        outs.append(Xd.cpu().numpy().T.copy())
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], Xh)
    for r in (0, 31, 32, 44):
        assert residual(A, outs[0][:, r], B[:, r]) < 1e-12
    F.close()


def test_bitwise_reproducible(smslu, W):
    A = W.laplacian_2d(150)
    n = A.shape[0]
    b = W.rhs(n, 1)
    outs = []
    for _ in range(2):
        F = smslu.ParallelSparseLU(A)
        x = np.empty(n); smslu.ldiv_(x, F, b)
        outs.append((x.copy(), F.L.data.copy()))
        smslu.lu_(F, A); smslu.ldiv_(x, F, b)
        outs.append((x.copy(), F.L.data.copy()))
        F.close()
    for x, l in outs[1:]:
        assert np.array_equal(x, outs[0][0]) and np.array_equal(l, outs[0][1])


@pytest.mark.parametrize("cfg", ["config2_lap2d_1024", "config5_lap3d_48"])
def test_full_size_properties(smslu, W, cfg):
    """BASELINE full sizes, where the oracle would take minutes: size-independent properties."""
    A = W.laplacian_2d(1024) if cfg.startswith("config2") else W.laplacian_3d(48)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A)
    st = F.stats()
    assert st["bad_pivot_col"] == -1
    x0 = W.rhs(n, 5) - 0.5
    b = A @ x0
    x = np.empty(n); smslu.ldiv_(x, F, b)
    assert residual(A, x, b) < 1e-13                                 # residual
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) < 1e-9        # round trip A*x0 -> x0
    b2 = W.rhs(n, 6)
    x2 = np.empty(n); smslu.ldiv_(x2, F, b2)
    x3 = np.empty(n); smslu.ldiv_(x3, F, 2.0 * b + b2)               # linearity
    assert np.linalg.norm(x3 - (2.0 * x + x2)) / np.linalg.norm(x3) < 1e-12
    y = b2.copy(); smslu.lsolve_(F, y); smslu.rsolve_(F, y)          # ldiv! == unperm(rsolve(lsolve(perm)))
    w = (F.Rs * b2)[F.p]
    smslu.lsolve_(F, w); smslu.rsolve_(F, w)
    xx = np.empty(n); xx[F.q] = w
    assert np.array_equal(xx, x2)
    # matrix right-hand sides at full size: 11 columns = one 8-wide and one 4-wide sweep through the bulk-level
    # variants of the solve kernels (256-row forward tiles, two-CTAs-per-SM backward), column by column
    # against the single-vector solve
    B = np.asfortranarray(np.stack([b2 * (1.0 + 0.1 * r) + r * b for r in range(11)], axis=1))
    X = np.empty((n, 11), order="F")
    smslu.ldiv_(X, F, B)
    for r in (0, 4, 7, 8, 10):
        xr = np.empty(n); smslu.ldiv_(xr, F, np.ascontiguousarray(B[:, r]))
        # a single right-hand side runs the chains of fronts in the persistent chain kernels, a block of them level by
        # level: two backward-stable sweeps with different summation orders agree to cond(A) * eps, and both solve A x = b
        assert np.linalg.norm(X[:, r] - xr) <= 1e-9 * np.linalg.norm(xr)
        assert residual(A, X[:, r], B[:, r]) < 1e-10 and residual(A, xr, B[:, r]) < 1e-10
    # refactor with shifted values (config 2: A + k*1e-3*I), pattern fixed
    A2 = sp.csc_matrix(A + 1e-3 * sp.identity(n)); A2.sort_indices()
    smslu.lu_(F, A2)
    smslu.ldiv_(x, F, b)
    assert residual(A2, x, b) < 1e-13
    F.close()


# ---------------------------------------------------------------------------------------------------------------
# Pivoting robustness: threshold test on the GPU, re-analysis with fresh host pivots

@pytest.mark.parametrize("n", [1, 2, 5, 8, 33, 100, 200])
def test_reference_dense_update_testset_default_arguments(smslu, W, n):
    """The reference's 'dense matrix' testset (test:108-146) with DEFAULT arguments: factorize rand(n,n), solve, new
    right-hand side, lu! with a NEW random matrix of the same (full) pattern, solve twice more -- each against an
    independent dense solve at the reference's tolerance (rtol = atol = 1e-10, test:26).  Static diagonal pivots are
    unstable on these matrices: the threshold test must catch it and the pivots must come from the host search."""
    A = W.dense_random(n, seed=2000 + n)
    F = smslu.ParallelSparseLU(A)
    for rnd, seed in ((0, 7), (0, 8)):
        b = W.rhs(n, seed)
        x = np.empty(n); smslu.ldiv_(x, F, b)
        assert isapprox(x, np.linalg.solve(A.toarray(), b), DENSE_TOL)
    A2 = W.dense_random(n, seed=3000 + n)
    assert smslu.lu_(F, A2) is None                                         # test:129-131
    for seed in (9, 10):
        b = W.rhs(n, seed)
        x = np.empty(n); smslu.ldiv_(x, F, b)
        assert isapprox(x, np.linalg.solve(A2.toarray(), b), DENSE_TOL)     # test:138, 144
    L, U = F.L, F.U
    assert abs(L).max() <= 1000.0 * (1 + 1e-6)                              # every multiplier passes the threshold
    B = (sp.diags(F.Rs) @ A2).tocsr()[F.p][:, F.q]
    assert abs(L @ U - B).max() < 1e-12 * max(1.0, abs(B).max())            # src:307
    F.close()


def test_threshold_violation_is_reported_not_silent(smslu, W):
    """New values that break the old pivots: strict ('native') objects return the distinct 'needs re-analysis' error
    instead of a wrong x; default objects re-pivot like the reference's lu! (src:245-279)."""
    A = W.laplacian_2d(12)
    n = A.shape[0]
    bad = A.copy()
    d = np.flatnonzero(bad.indices == np.repeat(np.arange(n), np.diff(bad.indptr)))
    bad.data[d[0]] = 1e-9                                                   # a tiny FIRST pivot (natural ordering below)
    b = W.rhs(n, 3)
    xo = np.linalg.solve(bad.toarray(), b)
    F = smslu.ParallelSparseLU(A, pivots="native", ordering="natural")
    with pytest.raises(smslu.PivotThresholdError):
        smslu.lu_(F, bad)
    assert F.stats()["threshold_col"] >= 0
    smslu.lu_(F, A)                                                         # the object stays usable with good values
    assert F.stats()["threshold_col"] == -1
    x = np.empty(n); smslu.ldiv_(x, F, b)
    assert residual(A, x, b) < 1e-14
    F.close()
    with pytest.raises(smslu.PivotThresholdError):
        smslu.ParallelSparseLU(bad, pivots="native", ordering="natural")
    G = smslu.ParallelSparseLU(A, ordering="natural")                       # default: re-pivot on the host when needed
    p_before = G.p.copy()
    smslu.lu_(G, bad)
    x = np.empty(n); smslu.ldiv_(x, G, b)
    assert isapprox(x, xo, 1e-10)
    assert not np.array_equal(G.p, p_before) or not np.array_equal(G.p, G.q)
    assert abs(G.L).max() <= 1000.0 * (1 + 1e-6)
    smslu.lu_(G, A)                                                         # and back: the new pivots still pass on A? either way correct
    smslu.ldiv_(x, G, b)
    assert isapprox(x, np.linalg.solve(A.toarray(), b), 1e-10)
    G.close()
    H = smslu.ParallelSparseLU(A, pivot_tol=-1.0, pivots="native", ordering="natural")   # test switched off: no error (and no guarantee)
    smslu.lu_(H, bad)
    H.close()


# ---------------------------------------------------------------------------------------------------------------
# Independent factorization at real sizes: SciPy SuperLU forced to the same pivots (F.p, F.q, F.Rs)

def _superlu_same_pivots(A, p, q, Rs):
    import scipy.sparse.linalg as spla
    B = sp.csc_matrix((sp.diags(Rs) @ A).tocsr()[p][:, q])
    lu = spla.splu(B, permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    assert np.array_equal(lu.perm_r, np.arange(A.shape[0])) and np.array_equal(lu.perm_c, np.arange(A.shape[0]))
    return lu


FULL = {
    "config2_lap2d_1024": lambda W: W.laplacian_2d(1024),                                    # BASELINE configs[1]
    "config3_lap3d_48": lambda W: W.laplacian_3d(48),
    "config4_blocks_nel45_x8": lambda W: W.block_border(nblocks=8, nel=45, ngr=5, border=64),  # full-size blocks
}


@pytest.mark.parametrize("cfg", list(FULL))
def test_full_size_entries_against_superlu_same_pivots(smslu, W, cfg):
    """L and U entry by entry against an independent LU (SuperLU, same pivot order) at BASELINE sizes."""
    A = FULL[cfg](W)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A)
    p, q, Rs = F.p, F.q, F.Rs
    lu = _superlu_same_pivots(A, p, q, Rs)
    Ls, Us = sp.csc_matrix(lu.L), sp.csc_matrix(lu.U)
    Ls.sort_indices(); Us.sort_indices()
    L, U = F.L, F.U
    # SuperLU's supernodes may store explicit zeros / drop nothing: compare on the union through sparse differences
    dl, du = abs(L - Ls), abs(U - Us)
    assert dl.max() < 1e-12 * max(1.0, abs(Ls).max()) and du.max() < 1e-12 * abs(Us).max()
    assert L.nnz == F.stats()["nnz_l_exact"] and (abs(Ls) > 0).nnz <= L.nnz
    if cfg.startswith("config2") or cfg.startswith("config3"):     # M-matrices: no cancellation -> plain relative error
        for M, Ms in ((L, Ls), (U, Us)):
            Ms = Ms.copy(); Ms.eliminate_zeros()                   # SuperLU pads its relaxed supernodes with zeros
            assert np.array_equal(M.indptr, Ms.indptr) and np.array_equal(M.indices, Ms.indices)   # bit-exact structure
            assert (abs(M.data - Ms.data) <= 1e-12 * abs(Ms.data)).all()
    b = W.rhs(n, 11)
    x = np.empty(n); smslu.ldiv_(x, F, b)
    w = (Rs * b)[p]
    xs = np.empty(n); xs[q] = lu.solve(w)
    assert np.linalg.norm(x - xs) <= 1e-10 * np.linalg.norm(xs)
    assert residual(A, x, b) <= max(4 * residual(A, xs, b), 1e-15)
    F.close()


def test_many_rhs_against_superlu_config5(smslu, W):
    """BASELINE config 5 shape: 64 right-hand sides on a 48^3 Laplacian, every column against SuperLU's solve."""
    A = W.laplacian_3d(48)
    n, nrhs = A.shape[0], 64
    F = smslu.ParallelSparseLU(A)
    lu = _superlu_same_pivots(A, F.p, F.q, F.Rs)
    B = W.rhs(n, 47, nrhs=nrhs)
    X = np.empty((n, nrhs), order="F")
    smslu.ldiv_(X, F, B)
    Wm = (F.Rs[:, None] * B)[F.p]
    Xs = np.empty_like(X); Xs[F.q] = lu.solve(np.ascontiguousarray(Wm))
    err = np.linalg.norm(X - Xs, axis=0) / np.linalg.norm(Xs, axis=0)
    assert err.max() < 1e-10
    F.close()


def test_north_star_lap3d_128_against_poisson_solver(smslu, W):
    """The north-star target at FULL size (3D 7-point Laplacian 128^3) on one GPU: refactorize + solve against an
    independent exact solver (type-I sine transform diagonalises the Dirichlet Laplacian) -- the checker bench.py
    prints as parity.x_relerr at every N."""
    import scipy.fft as sfft
    e = 128
    A = W.laplacian_3d(e)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A)
    assert F.stats()["bad_pivot_col"] == -1
    lam1 = 2.0 - 2.0 * np.cos(np.arange(1, e + 1) * np.pi / (e + 1))
    lam = lam1.reshape(-1, 1, 1) + lam1.reshape(1, -1, 1) + lam1.reshape(1, 1, -1)
    for k in (0, 1):
        vals = A.data.copy()
        if k:
            d = np.flatnonzero(A.indices == np.repeat(np.arange(n), np.diff(A.indptr)))
            vals[d] += 1e-3
            smslu.lu_(F, vals)                                    # refactorize: A + 1e-3 I, same pattern
        b = W.rhs(n, 47 + k)
        x = np.empty(n); smslu.ldiv_(x, F, b)
        xp = sfft.idstn(sfft.dstn(b.reshape(e, e, e), type=1, norm="ortho") / (lam + k * 1e-3), type=1, norm="ortho").reshape(-1)
        assert np.linalg.norm(x - xp) <= 1e-10 * np.linalg.norm(xp)
        Ak = A + k * 1e-3 * sp.identity(n)
        assert residual(Ak, x, b) < 1e-11                         # cond(A) ~ 7e3; the checker's own residual is 5e-13
    F.close()


def test_cuda_factors_through_the_reference_chunk_solve_config1(smslu, O, W):
    """BASELINE configs[0] (100 x 100): the CUDA factors fed to the restated reference solve -- dense column chunks,
    trsv + gemv per chunk, src:101-243 and src:349-392 -- give the same x as smslu.ldiv_ (and the same lsolve!/rsolve!)."""
    A = W.laplacian_2d(100)
    n = A.shape[0]
    F = smslu.ParallelSparseLU(A)
    L, U = F.L, F.U
    assert O.RefChunks.predict_bytes(L, U) < 6e9
    RC = O.RefChunks(L, U)
    b = W.rhs(n, 47)
    x = np.empty(n); smslu.ldiv_(x, F, b)
    xr = RC.ldiv(F.p, F.q, F.Rs, b)
    assert isapprox(x, xr, TOL)
    y = b.copy(); smslu.lsolve_(F, y)
    assert isapprox(y, RC.lsolve(b), TOL)
    y = b.copy(); smslu.rsolve_(F, y)
    assert isapprox(y, RC.rsolve(b), DENSE_TOL)
    RC.close(); F.close()
