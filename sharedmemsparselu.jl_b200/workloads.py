"""Deterministic synthetic inputs for the hot path.

Everything is generated from splitmix64 so that Python, C++ and the Julia shim can produce
bit-identical matrices and right-hand sides (the reference's own tests use Julia's
MersenneTwister(47), test/runtests.jl:35, which cannot be reproduced outside Julia).
All matrices are scipy CSC, Float64 values, int64 indices, natural ordering
idx = i + nx*(j + ny*k).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

_M64 = (1 << 64) - 1


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n uniform doubles in [0,1): element i is splitmix64 of state seed + (i+1)*golden."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & _M64) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _as_csc64(A) -> sp.csc_matrix:
    A = sp.csc_matrix(A)
    A.sort_indices()
    A.indptr = A.indptr.astype(np.int64)
    A.indices = A.indices.astype(np.int64)
    A.data = A.data.astype(np.float64)
    return A


def _lap1d(n: int) -> sp.spmatrix:
    return sp.diags([-np.ones(n - 1), -np.ones(n - 1)], [-1, 1], shape=(n, n), format="csr")


def laplacian_2d(nx: int, ny: int | None = None, shift: float = 0.0) -> sp.csc_matrix:
    """5-point stencil, diag 4 (+shift), off-diagonals -1, Dirichlet truncation."""
    ny = nx if ny is None else ny
    A = sp.kron(sp.identity(ny), _lap1d(nx)) + sp.kron(_lap1d(ny), sp.identity(nx))
    A = A + (4.0 + shift) * sp.identity(nx * ny)
    return _as_csc64(A)


def laplacian_3d(nx: int, ny: int | None = None, nz: int | None = None, shift: float = 0.0) -> sp.csc_matrix:
    """7-point stencil, diag 6 (+shift), off-diagonals -1."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    Ix, Iy, Iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    A = (sp.kron(Iz, sp.kron(Iy, _lap1d(nx))) + sp.kron(Iz, sp.kron(_lap1d(ny), Ix))
         + sp.kron(_lap1d(nz), sp.kron(Iy, Ix)))
    A = A + (6.0 + shift) * sp.identity(nx * ny * nz)
    return _as_csc64(A)


def fe_test_matrix(nel: int, ngr: int = 5, seed: int = 47) -> sp.csc_matrix:
    """Restates the reference's test fixture test_matrix (test/runtests.jl:12-21): nel dense
    ngr x ngr random blocks chained along the diagonal, consecutive blocks sharing one corner
    entry, the later block OVERWRITING the shared corner (test:18); entries U[0,1)."""
    n = nel * (ngr - 1) + 1
    vals = splitmix64(seed, nel * ngr * ngr).reshape(nel, ngr, ngr)
    M = np.zeros((n, n))
    for e in range(nel):
        lo = e * (ngr - 1)
        M[lo:lo + ngr, lo:lo + ngr] = vals[e].T      # column-major fill like Julia's rand(ngr,ngr)
    return _as_csc64(sp.csc_matrix(M))


def dense_random(n: int, seed: int = 47) -> sp.csc_matrix:
    """rand(n,n) turned sparse (test/runtests.jl:41-42), column-major fill, entries U[0,1)."""
    return _as_csc64(sp.csc_matrix(splitmix64(seed, n * n).reshape(n, n).T))


def rhs(n: int, seed: int = 47, nrhs: int = 1) -> np.ndarray:
    """Right-hand side(s): column r comes from seed+r.  Shape (n,) for nrhs==1 else (n,nrhs), F-order."""
    if nrhs == 1:
        return splitmix64(seed, n)
    B = np.empty((n, nrhs), order="F")
    for r in range(nrhs):
        B[:, r] = splitmix64(seed + r, n)
    return B


def block_border(nblocks: int = 64, nel: int = 45, ngr: int = 5, border: int = 64,
                 refactor_k: int = 0) -> sp.csc_matrix:
    """BASELINE config 4 (moment_kinetics-style): `nblocks` diagonal blocks, each the 2-D
    tensor-product pattern kron(P,P) of the FE fixture pattern P (n_b = (nel*(ngr-1)+1)^2),
    values U[0,1) (seed 1000+b) with +40 on the diagonal, scaled by (1+0.01*refactor_k);
    plus `border` coupling columns/rows: border column/row j is dense over block j mod nblocks
    (seeds 2000+j / 3000+j); corner G = border*I + U[0,1) (seed 4000)."""
    P = (fe_test_matrix(nel, ngr, seed=1).toarray() != 0).astype(np.float64)
    PP = sp.csc_matrix(sp.kron(sp.csc_matrix(P), sp.csc_matrix(P)))
    PP.sort_indices()
    nb = PP.shape[0]
    blocks = []
    for b in range(nblocks):
        B = PP.copy()
        B.data = splitmix64(1000 + b, B.nnz) * (1.0 + 0.01 * refactor_k)
        B = B + 40.0 * sp.identity(nb)
        blocks.append(B)
    D = sp.block_diag(blocks, format="csc")
    n0 = nb * nblocks
    rows, cols, vals = [], [], []
    for j in range(border):
        b = j % nblocks
        idx = np.arange(b * nb, (b + 1) * nb)
        rows.append(idx); cols.append(np.full(nb, j)); vals.append(splitmix64(2000 + j, nb))
    E = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n0, border))
    rows, cols, vals = [], [], []
    for j in range(border):
        b = j % nblocks
        idx = np.arange(b * nb, (b + 1) * nb)
        cols.append(idx); rows.append(np.full(nb, j)); vals.append(splitmix64(3000 + j, nb))
    Fm = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(border, n0))
    G = sp.csc_matrix(splitmix64(4000, border * border).reshape(border, border).T) + float(border) * sp.identity(border)
    return _as_csc64(sp.bmat([[D, E], [Fm, G]], format="csc"))
