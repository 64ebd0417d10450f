#!/bin/bash
# multi-GPU verification on N GPUs of one box: parity tests (worlds up to N) + the partitioned bench line at N (128^3)
N=${1:-2}; T=${2:-round2c}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/${T}_pytest_mgpu_n$N.log 2>&1; echo "pytest mgpu rc=$?" ); tail -3 gpurun_out/${T}_pytest_mgpu_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${T}_bench_n$N.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "parity_x_relerr", "residual")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["phases"])
except Exception as e:
    print("no json", e)
PY
grep -v "^built\|Warning\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/${T}_bench_n$N.err | tail -6
