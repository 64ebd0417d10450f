"""Per-kind device times of one 32-wide solve sweep (tensor-pipe kernels) next to an 8-wide one, 3D Laplacian."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import smslu
from sharedmemsparselu_jl_b200 import workloads as W, _capi
size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
A = W.laplacian_3d(size); n = A.shape[0]
F = smslu.ParallelSparseLU(A)
for nrhs in (8, 32):
    Bh = W.rhs(n, 47, nrhs=nrhs).reshape(n, nrhs, order="F")
    B = torch.from_numpy(np.ascontiguousarray(Bh.T)).cuda(); X = torch.empty_like(B)
    def solve():
        _capi.check(F._h, _capi.lib().smslu_solve(F._h, C.c_void_p(X.data_ptr()), n, C.c_void_p(B.data_ptr()), n, nrhs, n, n))
    solve(); solve()
    ms = F.stats()["ms_solve"]
    F.set_profile(True); solve(); sp = F.stats(); F.set_profile(False)
    print("nrhs %d: %.2f ms | " % (nrhs, ms) + "  ".join("%s %.2f (%d)" % (k, t, sp["launches_kernel"][k]) for k, t in sorted(sp["ms_kernel"].items(), key=lambda kv: -kv[1]) if t > 0.01))
os.environ["SMSLU_LEVEL_TIMES"] = "1"
F2 = smslu.ParallelSparseLU(A)
nrhs = 32
Bh = W.rhs(n, 47, nrhs=nrhs).reshape(n, nrhs, order="F")
B = torch.from_numpy(np.ascontiguousarray(Bh.T)).cuda(); X = torch.empty_like(B)
for rep in range(2):
    sys.stderr.write("==== rep %d\n" % rep)
    _capi.check(F2._h, _capi.lib().smslu_solve(F2._h, C.c_void_p(X.data_ptr()), n, C.c_void_p(B.data_ptr()), n, nrhs, n, n))
F.close(); F2.close()
