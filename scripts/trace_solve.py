"""Phase timestamps of CTA 0 of the last k_bwd launch (library built with SMSLU_TRACE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smslu
from sharedmemsparselu_jl_b200 import workloads as W, _capi
A = W.laplacian_2d(1024)
n = A.shape[0]
F = smslu.ParallelSparseLU(A)
b = W.rhs(n, 47); x = np.empty(n)
for rep in range(3):
    smslu.ldiv_(x, F, b)
    t = np.zeros(32, np.int64)
    _capi.lib().smslu_debug_trace(t.ctypes.data)
    r = t[16:24]
    print("k_bwd (last 1-CTA launch) load|wait|gather|gemv|..|diag:", [int(r[i] - r[0]) for i in range(6)])
    r = t[24:32]
    print("k_fwd (last 1-CTA launch) load|wait|gather..|diag|gemv:", [int(r[i] - r[0]) for i in range(5)])
F.close()
