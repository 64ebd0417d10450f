"""Run a BASELINE config through the CUDA path on one GPU: times (device, CUDA events), residual, properties.
usage: python scripts/configs_run.py lap3d 96 | bb 64 [nel] | lap2d 1024   [nrhs]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import smslu
from sharedmemsparselu_jl_b200 import workloads as W
which = sys.argv[1]; size = int(sys.argv[2])
if which == "lap3d": A = W.laplacian_3d(size)
elif which == "lap2d": A = W.laplacian_2d(size)
else: A = W.block_border(nblocks=size, nel=int(sys.argv[3]) if len(sys.argv) > 3 else 45)
n = A.shape[0]
t = time.time(); F = smslu.ParallelSparseLU(A); t_first = time.time() - t
st = F.stats()
b = W.rhs(n, 47); x = np.empty(n)
smslu.ldiv_(x, F, b)
res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
times = []
for k in range(3):
    A2 = A.copy(); A2.data = A.data * (1.0 + 0.01 * k)
    smslu.lu_(F, A2); smslu.ldiv_(x, F, b)
    s2 = F.stats(); times.append((s2["ms_refactor"], s2["ms_solve"]))
res2 = np.linalg.norm(A2 @ x - b) / np.linalg.norm(b)
rf = min(t[0] for t in times); sv = min(t[1] for t in times)
print("%s %d: n=%d nnzA=%d nnzL=%.3e flops=%.3e levels=%d maxfront=%d | analyze %.1fs first-call %.1fs | refactor %.2f ms (%.2f TFLOP/s) solve %.2f ms (%.0f GB/s of 12 B/nnz) launches %d+%d | residual %.1e / %.1e" % (
    which, size, n, A.nnz, st["nnz_l_exact"], st["flops_exact"], st["n_levels"], st["max_front"], st["ms_analyze"] / 1e3, t_first,
    rf, st["flops_exact"] / rf / 1e9, sv, 24.0 * st["nnz_l_exact"] / sv / 1e6, s2["launches_refactor"], s2["launches_solve"], res, res2), flush=True)
F.close()
