#!/bin/bash
# A/B several builds of libsmslu.so on the same box: gpurun_ab/libsmslu_<V>.so; usage: abn.sh "A B C" [lap3d edge] [reps]
VARS=${1:-"A B"}; EDGE=${2:-128}; REPS=${3:-2}
for rep in $(seq 1 $REPS); do for v in $VARS; do
SMSLU_LIB=$PWD/gpurun_ab/libsmslu_$v.so python - <<PY
import sys
sys.path.insert(0, ".")
import torch, numpy as np, smslu
from sharedmemsparselu_jl_b200 import workloads as W
A = W.laplacian_3d($EDGE); n = A.shape[0]
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
tr = timed(lambda: F.refactor_async(v), 3); ts = timed(lambda: F.solve_async(x, b), 5)
xh = x.cpu().numpy(); bh = b.cpu().numpy()
res = np.linalg.norm(A @ xh - bh) / np.linalg.norm(bh)
print("$v rep $rep: lap3d $EDGE refactor %.2f ms (%.2f TFLOP/s) solve %.3f ms residual %.2e" % (tr, F.stats()["flops_exact"] / tr / 1e9, ts, res), flush=True)
F.close()
PY
done; done
