"""Worker of tests/test_multigpu_host.py: one rank of a world_size-N gloo group walking the partitioned
factorization + solve on the CPU (tests/hostexec.cpp) with torch.distributed all-reduces at the three
exchange points, checked against the single-rank walk and the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import smslu  # noqa: E402,F401
from sharedmemsparselu_jl_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from conftest import HostExec, PartitionedWalk  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(a):
        if a.size:
            t = torch.from_numpy(a)
            dist.all_reduce(t)

    hx = HostExec()
    cases = {"lap2d_40": W.laplacian_2d(40), "lap3d_10": W.laplacian_3d(10),
             "block_border": W.block_border(nblocks=4, nel=5, ngr=5, border=8)}
    for name, A in cases.items():
        n = A.shape[0]
        Rs = O.row_scale_sum(A)
        P = PartitionedWalk(hx, A, world, rank, allreduce)
        own = P.owner
        assert P.info["ntop"] > 0 and P.info["lu_top_size"] > 0, name
        for r in range(world):
            assert np.any(own == r), (name, "rank without work", r)
        for k in range(2):                                    # factor, then refactor with new values
            Ax = A.data * (1.0 + 0.01 * k)
            bad = P.factor(Ax, Rs)
            assert bad == -1
            b = W.rhs(n, 47 + k)
            x = P.solve(b)
            A2 = A.copy(); A2.data = Ax
            ref = O.OracleLU(A2, p=P.p, q=P.q, Rs=Rs)
            xo = ref.solve(b)
            assert np.linalg.norm(x - xo) <= 1e-12 * np.linalg.norm(xo), (name, k)
            res = np.linalg.norm(A2 @ x - b) / np.linalg.norm(b)
            assert res <= max(4 * np.linalg.norm(A2 @ xo - b) / np.linalg.norm(b), 1e-15)
            # every rank ends with the same full solution
            t = torch.from_numpy(x.copy()); dist.broadcast(t, 0)
            assert np.array_equal(t.numpy(), x)
        P.close()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
