#!/bin/bash
# the partitioned bench line at N GPUs (128^3), short
N=${1:-8}; T=${2:-round2_final}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${T}_bench_n$N.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "parity_x_relerr", "residual")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["phases"], d["comm"]["ms_per_profiled_step"])
except Exception as e:
    print("no json", e)
PY
grep -v "^built\|Warning\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/${T}_bench_n$N.err | tail -6
