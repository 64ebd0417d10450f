"""Device time of refactor-only and solve-only loops on a 3D Laplacian (lanes on unless SMSLU_NO_LANES)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, smslu
from sharedmemsparselu_jl_b200 import workloads as W
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 96
A = W.laplacian_3d(edge); n = A.shape[0]
F = smslu.ParallelSparseLU(A)
st = torch.cuda.current_stream(); F.set_stream(st)
v = torch.from_numpy(A.data.copy()).cuda(); b = torch.from_numpy(W.rhs(n, 47)).cuda(); x = torch.empty_like(b)
def timed(fn, reps):
    fn(); F.sync(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); F.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("lap3d %d: solve only %.3f ms (launches %d)   refactor only %.2f ms" % (edge, timed(lambda: F.solve_async(x, b), 10), F.stats()["launches_solve"], timed(lambda: F.refactor_async(v), 2)))
F.close()
