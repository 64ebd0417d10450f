"""Per-level device time of one refactor + solve (SMSLU_LEVEL_TIMES=1 debug marks), 2D grid."""
import os, sys
os.environ["SMSLU_LEVEL_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, smslu
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
for rep in range(3):
    sys.stderr.write("==== rep %d\n" % rep)
    smslu.lu_(F, A)
    smslu.ldiv_(x, F, b)
F.close()
