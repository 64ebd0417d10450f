"""Per-kernel-kind device time of one refactor + solve (smslu_set_profile) for a given workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smslu
from sharedmemsparselu_jl_b200 import workloads as W
which = sys.argv[1]; size = int(sys.argv[2])
A = W.laplacian_3d(size) if which == "lap3d" else W.laplacian_2d(size)
n = A.shape[0]
F = smslu.ParallelSparseLU(A)
b = W.rhs(n, 47); x = np.empty(n)
F.set_profile(True)
smslu.lu_(F, A); smslu.ldiv_(x, F, b)
st = F.stats()
F.set_profile(False)
tot = sum(st["ms_kernel"].values())
print("%s %d: total %.2f ms (refactor %.2f solve %.2f) flops %.3e" % (which, size, tot, st["ms_refactor"], st["ms_solve"], st["flops_exact"]))
for k, v in sorted(st["ms_kernel"].items(), key=lambda kv: -kv[1]):
    if v > 0: print("  %-14s %9.3f ms  %5.1f%%  launches %d" % (k, v, 100 * v / tot, st["launches_kernel"][k]))
F.close()
