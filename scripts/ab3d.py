"""A/B helper: refactor-only / solve-only device times and the per-kind profile of one library build on a 3D Laplacian.
usage: SMSLU_LIB=path/to/libsmslu_X.so python scripts/ab3d.py [edge] [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, smslu
from sharedmemsparselu_jl_b200 import workloads as W
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 128
tag = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(os.environ.get("SMSLU_LIB", "default"))
A = W.laplacian_3d(edge); n = A.shape[0]
F = smslu.ParallelSparseLU(A)
st = torch.cuda.current_stream(); F.set_stream(st)
v = torch.from_numpy(A.data.copy()).cuda(); bh = W.rhs(n, 47); b = torch.from_numpy(bh).cuda(); x = torch.empty_like(b)
def timed(fn, reps):
    fn(); F.sync(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); F.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ts = timed(lambda: F.solve_async(x, b), 10)
tr = timed(lambda: F.refactor_async(v), 3)
F.solve_async(x, b); F.sync()
xh = x.cpu().numpy()
res = np.linalg.norm(A @ xh - bh) / np.linalg.norm(bh)
F.set_profile(True)
F.refactor_async(v); F.solve_async(x, b); F.sync()
sp = F.stats(); F.set_profile(False)
kinds = "  ".join("%s %.2f" % (k, t) for k, t in sorted(sp["ms_kernel"].items(), key=lambda kv: -kv[1]) if t > 0.05)
print("%s lap3d %d: refactor %.2f ms  solve %.3f ms  residual %.2e | %s" % (tag, edge, tr, ts, res, kinds))
F.close()
