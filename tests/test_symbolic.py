"""CPU: host-side symbolic analysis of the product (csrc/symbolic.cpp) against the oracle.

Bit-exact structure: the supernodal layout, walked on the CPU by tests/hostexec.cpp exactly the way
the GPU kernels walk it, must reproduce the oracle's L/U pattern entry for entry and its values to
1e-12 relative when both use the same (p, q, Rs)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import relerr


def cases(W):
    return {
        "lap2d_10": (W.laplacian_2d(10), {}),
        "lap2d_37x23": (W.laplacian_2d(37, 23), {}),
        "lap2d_37x23_grid": (W.laplacian_2d(37, 23), dict(grid=(37, 23, 1), ordering=4)),
        "lap2d_64_norelax": (W.laplacian_2d(64), dict(relax=0)),
        "lap2d_64_w8": (W.laplacian_2d(64), dict(maxw=8)),
        "lap2d_100": (W.laplacian_2d(100), {}),
        "lap3d_12": (W.laplacian_3d(12), {}),
        "lap3d_9x7x5_grid": (W.laplacian_3d(9, 7, 5), dict(grid=(9, 7, 5), ordering=4)),
        "natural": (W.laplacian_2d(15), dict(ordering=1)),
        "fe_50": (W.fe_test_matrix(50, seed=3) + 5 * sp.identity(201), {}),
        "dense_40": (W.dense_random(40, seed=2) + 40 * sp.identity(40), {}),
        "n1": (sp.csc_matrix(np.array([[2.5]])), {}),
        "diag": (sp.identity(7, format="csc") * 3.0, {}),
    }


@pytest.mark.parametrize("name", ["lap2d_10", "lap2d_37x23", "lap2d_37x23_grid", "lap2d_64_norelax", "lap2d_64_w8",
                                  "lap2d_100", "lap3d_12", "lap3d_9x7x5_grid", "natural", "fe_50", "dense_40",
                                  "n1", "diag"])
def test_layout_reproduces_oracle(hostexec, O, W, name):
    A, kw = cases(W)[name]
    A = sp.csc_matrix(A)
    n = A.shape[0]
    Rs = O.row_scale_sum(A)
    r = hostexec.run(A, Rs=Rs, **kw)
    assert r["bad"] == -1
    p, q = r["p"], r["q"]
    assert np.array_equal(np.sort(p), np.arange(n)) and np.array_equal(p, q)
    F = O.OracleLU(A, p=p, q=q, Rs=Rs)
    # bit-exact symbolic structure
    assert r["nnzL_exact"] == F.L.nnz == F.U.nnz
    assert np.array_equal(r["Lp"], F.Lp) and np.array_equal(r["Li"], F.Li)
    assert np.array_equal(r["Up"], F.Up) and np.array_equal(r["Ui"], F.Ui)
    # values: 1e-12 relative (BASELINE north_star)
    assert relerr(r["Lx"], F.Lx) < 1e-12
    assert relerr(r["Ux"], F.Ux) < 1e-12
    # solves through the update-vector scheme the kernels use
    assert np.allclose(r["x"], F.solve(r["b"]), rtol=1e-11, atol=1e-13)
    assert np.allclose(r["lsolve_b"], F.lsolve(r["b"]), rtol=1e-12, atol=1e-14)
    assert np.allclose(r["rsolve_b"], F.usolve(r["b"]), rtol=1e-10, atol=1e-13)
    res = np.linalg.norm(A @ r["x"] - r["b"]) / np.linalg.norm(r["b"])
    res_o = np.linalg.norm(A @ F.solve(r["b"]) - r["b"]) / np.linalg.norm(r["b"])
    assert res <= max(4 * res_o, 1e-14)


def test_given_unsymmetric_permutations(hostexec, O, W):
    """ordering GIVEN with p != q (what the Julia shim passes from UMFPACK): contract still holds."""
    A = sp.csc_matrix(W.dense_random(25, seed=9))
    Rs = O.row_scale_sum(A)
    F0 = O.OracleLU(A, Rs=Rs, diag_tol=2.0)        # classical partial pivoting picks p
    assert not np.array_equal(F0.p, F0.q)
    r = hostexec.run(A, ordering=2, p=F0.p, q=F0.q, Rs=Rs)
    # the layout may compose (p,q) with a postorder; the contract is what matters
    F = O.OracleLU(A, p=r["p"], q=r["q"], Rs=Rs)
    assert np.array_equal(r["Li"], F.Li) and np.array_equal(r["Ui"], F.Ui)
    assert relerr(r["Lx"], F.Lx) < 1e-11 and relerr(r["Ux"], F.Ux) < 1e-11
    assert np.allclose(A @ r["x"], r["b"], rtol=1e-10, atol=1e-10)


def test_unsymmetric_pattern_is_padded(hostexec, O, W):
    """A pattern that is not symmetric gets the pattern of A+A' (explicit zeros)."""
    A = sp.csc_matrix(sp.triu(W.laplacian_2d(8), 0) + sp.tril(W.laplacian_2d(8), -3))
    r = hostexec.run(A, ordering=1)
    As = A + 0 * A.T
    Apad = sp.csc_matrix(A + sp.csc_matrix((np.zeros(A.T.nnz), A.T.tocsc().indices, A.T.tocsc().indptr), shape=A.shape))
    F = O.OracleLU(A, p=r["p"], q=r["q"])
    assert np.allclose(A @ r["x"], r["b"], rtol=1e-12, atol=1e-12)
    assert r["nnzL_exact"] >= F.L.nnz


def test_library_symbolic_matches_hostexec(smslu, hostexec, W):
    """The shared library's analysis (through the C ABI) is the same code path as hostexec's."""
    from sharedmemsparselu_jl_b200 import _SymbolicOnly
    A = W.laplacian_2d(40)
    F = _SymbolicOnly(A)
    r = hostexec.run(A, factor=False)
    st = F.stats()
    assert np.array_equal(F.p, r["p"]) and np.array_equal(F.q, r["q"])
    assert st["n_supernodes"] == r["nsn"] and st["n_levels"] == r["nlevels"]
    assert st["nnz_l_exact"] == r["nnzL_exact"] and st["lu_pool_doubles"] == r["lu_size"]
    sym = F.symbolic()
    sn = sym["sn_start"]
    assert sn[0] == 0 and sn[-1] == A.shape[0] and np.all(np.diff(sn) > 0) and np.all(np.diff(sn) <= 128)
    # every supernode's row list is sorted, beyond its last column, and matches the column count
    for s in range(st["n_supernodes"]):
        rows = sym["rows"][sym["rows_ptr"][s]:sym["rows_ptr"][s + 1]]
        assert np.all(np.diff(rows) > 0)
        if rows.size:
            assert rows[0] >= sn[s + 1]
        assert rows.size == sym["colcount"][sn[s + 1] - 1] - 1
    # parents come later, levels increase towards the root
    par = sym["sn_parent"]
    for s in range(st["n_supernodes"]):
        if par[s] >= 0:
            assert par[s] > s and sym["sn_level"][par[s]] > sym["sn_level"][s]
    F.close()


def test_orderings_reduce_fill(hostexec, W):
    A = W.laplacian_2d(48)
    nat = hostexec.run(A, ordering=1, factor=False)["nnzL_exact"]
    ndg = hostexec.run(A, ordering=3, factor=False)["nnzL_exact"]
    grd = hostexec.run(A, ordering=4, grid=(48, 48, 1), factor=False)["nnzL_exact"]
    assert ndg < 0.6 * nat and grd < 0.7 * nat


def test_block_border_components_and_dense_rows(hostexec, O, W):
    """Config-4 style matrix (small): independent diagonal blocks + dense border."""
    A = W.block_border(nblocks=4, nel=3, ngr=4, border=4)
    r = hostexec.run(A)
    assert r["bad"] == -1
    F = O.OracleLU(A, p=r["p"], q=r["q"])
    assert np.array_equal(r["Li"], F.Li)
    assert relerr(r["Lx"], F.Lx) < 1e-11
    assert np.allclose(A @ r["x"], r["b"], rtol=1e-11, atol=1e-11)


def test_analysis_does_not_depend_on_host_threads(smslu, W, monkeypatch):
    """The ordering, the graph construction and the scatter map run on several host threads above a size
    threshold (n >= 100 000, >= 2e6 entries); the layout must be identical for any thread count."""
    from sharedmemsparselu_jl_b200 import _SymbolicOnly
    A = W.laplacian_2d(720)                      # n = 518 400, 2.6e6 entries: every threaded path is taken
    outs = []
    for nthreads in ("1", "6", "3"):
        monkeypatch.setenv("SMSLU_HOST_THREADS", nthreads)
        F = _SymbolicOnly(A)
        sym = F.symbolic()
        st = F.stats()
        outs.append((F.p.copy(), F.q.copy(), sym, {k: st[k] for k in ("n_supernodes", "n_levels", "nnz_l_exact",
                                                                      "lu_pool_doubles", "cb_pool_doubles", "flops_exact")}))
        F.close()
    p0, q0, sym0, st0 = outs[0]
    for p, q, sym, st in outs[1:]:
        assert np.array_equal(p, p0) and np.array_equal(q, q0)
        assert st == st0
        for k in sym0:
            assert np.array_equal(sym[k], sym0[k]), k
