"""Host-side probe of the BASELINE configs: analysis time, fill, storage, levels (no GPU needed)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smslu
from sharedmemsparselu_jl_b200 import _SymbolicOnly, workloads as W
which = sys.argv[1]; size = int(sys.argv[2])
t = time.time()
if which == "lap3d": A = W.laplacian_3d(size)
elif which == "lap2d": A = W.laplacian_2d(size)
else: A = W.block_border(nblocks=size, nel=int(sys.argv[3]) if len(sys.argv) > 3 else 45)
print("matrix n=%d nnz=%d built in %.1fs" % (A.shape[0], A.nnz, time.time() - t))
t = time.time(); S = _SymbolicOnly(A); ta = time.time() - t
st = S.stats()
print("analyze %.1fs nnzL=%.3e stored=%.3e flops=%.3e nsn=%d levels=%d maxfront=%d lu_pool=%.2f GB cb_pool=%.2f GB" % (
    ta, st["nnz_l_exact"], st["nnz_l_stored"], st["flops_exact"], st["n_supernodes"], st["n_levels"], st["max_front"],
    8e-9 * st["lu_pool_doubles"], 8e-9 * st["cb_pool_doubles"]))
