"""Phase timestamps of the last panel launch (library built with SMSLU_TRACE=1)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smslu
from sharedmemsparselu_jl_b200 import workloads as W, _capi
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = W.laplacian_2d(grid)
F = smslu.ParallelSparseLU(A)
for rep in range(3):
    smslu.lu_(F, A)
    t = np.zeros(32, np.int64)
    n = _capi.lib().smslu_debug_trace(t.ctypes.data)
    r, p = t[:8], t[8:16]
    print("row warp  :", [int(r[i] - r[0]) for i in range(8)])
    print("pivot warp:", [int(p[i] - p[0]) for i in range(7)])
F.close()
