"""GPU probe: factor + solve a few sizes, print timing / parity summaries (development tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smslu
from sharedmemsparselu_jl_b200 import workloads as W

def run(name, A, reps=3, **kw):
    n = A.shape[0]
    t = time.time(); F = smslu.ParallelSparseLU(A, **kw); t_first = time.time() - t
    st = F.stats()
    b = W.rhs(n, 47); x = np.empty(n)
    best_f, best_s = 1e30, 1e30
    for _ in range(reps):
        smslu.lu_(F, A); smslu.ldiv_(x, F, b)
        s2 = F.stats(); best_f = min(best_f, s2["ms_refactor"]); best_s = min(best_s, s2["ms_solve"])
    res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    s2 = F.stats()
    print(f"{name}: n={n} nnzL={st['nnz_l_exact']/1e6:.2f}M stored={st['nnz_l_stored']/1e6:.2f}M nsn={st['n_supernodes']} "
          f"levels={st['n_levels']} flops={st['flops_exact']:.3g} analyze={st['ms_analyze']:.0f}ms upload={st['ms_upload']:.0f}ms "
          f"first={t_first:.2f}s refactor={best_f:.3f}ms ({st['flops_stored']/best_f/1e9:.2f} TF/s stored) solve={best_s:.3f}ms "
          f"({(24*st['nnz_l_exact'])/best_s/1e6:.1f} GB/s) launches={s2['launches_refactor']}+{s2['launches_solve']} "
          f"h2d={s2['ms_refactor_h2d']:.3f}ms res={res:.2e}", flush=True)
    F.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["s", "m", "l"]
    if "s" in which:
        run("lap2d_100", W.laplacian_2d(100))
        run("lap3d_20", W.laplacian_3d(20))
    if "m" in which:
        run("lap2d_256", W.laplacian_2d(256))
        run("lap2d_512", W.laplacian_2d(512))
        run("lap3d_32", W.laplacian_3d(32))
    if "l" in which:
        run("lap2d_1024", W.laplacian_2d(1024))
        run("lap3d_48", W.laplacian_3d(48))
        run("lap3d_64", W.laplacian_3d(64))
