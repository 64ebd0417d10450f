"""Partitioned run of a 3D Laplacian over the GPUs of one box (torchrun): refactor + solve times, residual."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("gloo")
import smslu
from sharedmemsparselu_jl_b200 import workloads as W
which = sys.argv[1]; size = int(sys.argv[2])
A = W.laplacian_3d(size) if which == "lap3d" else W.laplacian_2d(size)
n = A.shape[0]
ids = [smslu.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, 0)
t = time.time()
F = smslu.ParallelSparseLU(A, device=lrank, nranks=world, rank=rank, comm_id=ids[0])
t_first = time.time() - t
st = F.stats()
b = W.rhs(n, 47); x = np.empty(n)
smslu.ldiv_(x, F, b)
times = []
for k in range(3):
    A2 = A.copy(); A2.data = A.data * (1.0 + 0.01 * k)
    dist.barrier()
    t0 = time.perf_counter(); smslu.lu_(F, A2); t1 = time.perf_counter(); smslu.ldiv_(x, F, b); t2 = time.perf_counter()
    s2 = F.stats(); times.append((s2["ms_refactor"], s2["ms_solve"], (t1 - t0) * 1e3, (t2 - t1) * 1e3))
res = np.linalg.norm(A2 @ x - b) / np.linalg.norm(b)
tt = torch.tensor([min(t[0] for t in times), min(t[1] for t in times)], dtype=torch.float64)
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print("%s %d on %d GPUs: n=%d nnzL=%.3e flops=%.3e top supernodes %d (local %d) all-reduce %.1f MB/refactor | refactor %.2f ms (%.2f TFLOP/s aggregate) solve %.2f ms | wall refactor %.1f ms solve %.1f ms | residual %.1e" % (
        which, size, world, n, st["nnz_l_exact"], st["flops_exact"], st["n_top_supernodes"], st["n_local_supernodes"],
        8e-6 * st["allreduce_doubles_refactor"], tt[0], st["flops_exact"] / float(tt[0]) / 1e9, tt[1],
        min(t[2] for t in times), min(t[3] for t in times), res), flush=True)
F.close()
dist.barrier()
dist.destroy_process_group()
