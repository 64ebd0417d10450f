"""BASELINE config 5: many-RHS solve sweep on a 3D Laplacian (device time of smslu_solve per nrhs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import smslu
from sharedmemsparselu_jl_b200 import workloads as W
size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
A = W.laplacian_3d(size)
n = A.shape[0]
F = smslu.ParallelSparseLU(A)
st = F.stats()
print("lap3d %d: n=%d nnzL=%.3e refactor %.1f ms" % (size, n, st["nnz_l_exact"], st["ms_refactor"]))
for nrhs in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    Bh = W.rhs(n, 47, nrhs=nrhs).reshape(n, nrhs, order="F")
    B = torch.from_numpy(np.ascontiguousarray(Bh.T)).cuda()      # (nrhs, n) row-major == (n, nrhs) column-major, ld = n
    X = torch.empty_like(B)
    from sharedmemsparselu_jl_b200 import _capi
    import ctypes as C
    for rep in range(2):
        _capi.check(F._h, _capi.lib().smslu_solve(F._h, C.c_void_p(X.data_ptr()), n, C.c_void_p(B.data_ptr()), n, nrhs, n, n))
    ms = F.stats()["ms_solve"]
    x0 = X[0].cpu().numpy(); b0 = B[0].cpu().numpy()
    res = np.linalg.norm(A @ x0 - b0) / np.linalg.norm(b0)
    xl = X[nrhs - 1].cpu().numpy(); bl = B[nrhs - 1].cpu().numpy()
    resl = np.linalg.norm(A @ xl - bl) / np.linalg.norm(bl)
    print("nrhs %3d: %8.2f ms  %7.3f ms/rhs  %7.1f solves/s  effective %6.0f GB/s (24 B/nnz per rhs)  residual %.1e / %.1e" % (
        nrhs, ms, ms / nrhs, nrhs / ms * 1e3, 24.0 * st["nnz_l_exact"] * nrhs / ms / 1e6, res, resl), flush=True)
F.close()
