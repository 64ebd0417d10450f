#!/bin/bash
# Round-2 evidence, call B (4 GPUs of one box): multi-GPU parity tests (world 2 and 4), partitioned bench lines N=2 and N=4 on 128^3
T=round2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
( timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/${T}_pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$?" ); tail -4 gpurun_out/${T}_pytest_mgpu.log
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
  echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_n$N.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "parity_x_relerr", "residual")}, d["e2e"], d["comm"], d["phases"])
except Exception as e:
    print("no json", e)
PY
  grep -v "^built\|Warning\|warn" gpurun_out/${T}_bench_n$N.err | tail -8
done
