#!/bin/bash
# A/B two builds of libsmslu.so on the same box: gpurun_ab/libsmslu_A.so vs gpurun_ab/libsmslu_B.so
for rep in 1 2 3; do for v in A B; do
SMSLU_LIB=$PWD/gpurun_ab/libsmslu_$v.so python - <<PY
import sys, json, subprocess
sys.path.insert(0, ".")
import torch, numpy as np, smslu
from sharedmemsparselu_jl_b200 import workloads as W
A = W.laplacian_2d(1024); n = A.shape[0]
F = smslu.ParallelSparseLU(A)
st = torch.cuda.current_stream(); F.set_stream(st)
v = torch.from_numpy(A.data.copy()).cuda(); b = torch.from_numpy(W.rhs(n, 47)).cuda(); x = torch.empty_like(b)
for _ in range(5): F.refactor_async(v); F.solve_async(x, b)
F.sync(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(20): F.refactor_async(v); F.solve_async(x, b)
e1.record(st); F.sync(); torch.cuda.synchronize()
xh = x.cpu().numpy(); bh = b.cpu().numpy()
res = np.abs(A @ xh - bh).max() / np.abs(bh).max()
print("$v rep $rep: %.3f ms/step  residual %.2e" % (e0.elapsed_time(e1) / 20, res))
F.close()
PY
done; done
