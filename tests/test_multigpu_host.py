"""CPU, world_size 2, 4 and 8 over gloo: the host-side logic of the multi-GPU path -- partition of the
elimination tree into per-rank subtrees + a replicated top, the three exchange points (top panels +
interface contribution blocks, forward-solve interface vectors, solution gather) -- walked by
tests/hostexec.cpp over exactly the layout the CUDA kernels consume."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_walk_over_gloo(world, hostexec, O):
    port = 29600 + world + (os.getpid() % 200)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_mgpu_host_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, out[-3000:])


def test_partition_properties(hostexec, W):
    """Subtrees are closed under 'child of', the top is closed under 'parent of', loads are balanced."""
    from conftest import PartitionedWalk
    A = W.laplacian_2d(64)
    for G in (2, 4, 8):
        P = PartitionedWalk(hostexec, A, G, 0, lambda a: None)
        own = P.owner
        assert set(np.unique(own)) == set(range(-1, G))
        assert P.info["ntop"] == int((own == -1).sum()) and 0 < P.info["ntop"] < own.size // 4
        P.close()
