"""One analysis + first factorization, then ONE refactorize+solve step: the command profiled by ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smslu
from sharedmemsparselu_jl_b200 import workloads as W

# usage: <script> [grid]  (2D)   or   <script> lap3d <edge>
if len(sys.argv) > 2 and sys.argv[1] == "lap3d":
    grid = int(sys.argv[2]); A = W.laplacian_3d(grid)
else:
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    A = W.laplacian_2d(grid)
n = A.shape[0]
F = smslu.ParallelSparseLU(A)
b = W.rhs(n, 47); x = np.empty(n)
smslu.lu_(F, A)
smslu.ldiv_(x, F, b)
st = F.stats()
print("one_step grid=%d refactor %.3f ms solve %.3f ms launches %d+%d residual %.2e" % (
    grid, st["ms_refactor"], st["ms_solve"], st["launches_refactor"], st["launches_solve"],
    np.linalg.norm(A @ x - b) / np.linalg.norm(b)))
F.close()
