"""One analysis + factorization, then ONE 32-wide solve sweep (tensor-pipe kernels): the command profiled by ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smslu
from sharedmemsparselu_jl_b200 import workloads as W
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 96
A = W.laplacian_3d(edge); n = A.shape[0]
F = smslu.ParallelSparseLU(A)
B = W.rhs(n, 47, nrhs=32); X = np.empty((n, 32), order="F")
smslu.ldiv_(X, F, B)
print("one_step_wide edge=%d solve(32 rhs) %.3f ms residual %.2e" % (edge, F.stats()["ms_solve"], np.linalg.norm(A @ X[:, 31] - B[:, 31]) / np.linalg.norm(B[:, 31])))
F.close()
