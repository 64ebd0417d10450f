"""Worker of tests/test_multigpu_gpu.py: one process per GPU; the partitioned CUDA path through the C ABI
(NCCL all-reduces inside libsmslu.so) against the CPU oracle."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import smslu  # noqa: E402
from sharedmemsparselu_jl_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from conftest import relerr_csc  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cases = {"lap2d_64": W.laplacian_2d(64), "lap3d_14": W.laplacian_3d(14), "lap2d_300": W.laplacian_2d(300),
             "block_border": W.block_border(nblocks=8, nel=6, ngr=5, border=8)}
    for name, A in cases.items():
        n = A.shape[0]
        ids = [smslu.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, 0)
        F = smslu.ParallelSparseLU(A, device=rank, nranks=world, rank=rank, comm_id=ids[0])
        st = F.stats()
        assert st["n_top_supernodes"] > 0 and st["n_local_supernodes"] > 0
        for k in range(2):
            A2 = sp.csc_matrix(A * (1.0 + 0.01 * k) + k * 1e-3 * sp.identity(n)); A2.sort_indices()
            if k:
                smslu.lu_(F, A2)
            b = W.rhs(n, 47 + k)
            x = np.empty(n)
            smslu.ldiv_(x, F, b)
            ref = O.OracleLU(A2, p=F.p, q=F.q, Rs=F.Rs)
            xo = ref.solve(b)
            assert np.linalg.norm(x - xo) <= 1e-12 * np.linalg.norm(xo), (name, k, np.linalg.norm(x - xo) / np.linalg.norm(xo))
            res = np.linalg.norm(A2 @ x - b) / np.linalg.norm(b)
            res_o = np.linalg.norm(A2 @ xo - b) / np.linalg.norm(b)
            assert res <= max(4 * res_o, 1e-15), (name, k, res, res_o)      # no worse than the oracle's
            # the ranks' shares of L and U sum to the oracle's factors, structure bit-exact
            L, U = F.L, F.U
            assert np.array_equal(L.indices, ref.Li) and np.array_equal(U.indices, ref.Ui)
            lx = torch.from_numpy(L.data.copy()); ux = torch.from_numpy(U.data.copy())
            dist.all_reduce(lx); dist.all_reduce(ux)
            tol = 1e-12 if name.startswith("lap") else 1e-9
            assert relerr_csc(lx.numpy(), ref.Lx, ref.Lp) < tol and relerr_csc(ux.numpy(), ref.Ux, ref.Up) < tol, name
            # lsolve!/rsolve! across the partition
            y = b.copy(); smslu.lsolve_(F, y)
            assert np.linalg.norm(y - ref.lsolve(b)) <= 1e-12 * np.linalg.norm(y)
            y = b.copy(); smslu.rsolve_(F, y)
            assert np.linalg.norm(y - ref.usolve(b)) <= 1e-10 * np.linalg.norm(y)
        F.close()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
