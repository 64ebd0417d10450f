#!/bin/bash
# usage: bash scripts/gpu_check_n.sh TAG N   -> multi-GPU pytest for world N and bench at N
TAG=${1:-x}; N=${2:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -k "[$N]" > gpurun_out/pytest_mgpu_${TAG}_n$N.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_mgpu_${TAG}_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "bench n$N rc=$?"; tail -3 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_n$N.json") if l.startswith("{")][-1])
    print("N=$N ms/step %.3f e2e %.3f res %.2e %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["residual"], d["scaling"]))
    print({k: v["ms"] for k, v in d["kernels"].items()})
    print(d["config"]["parallelism"])
except Exception as e:
    print("bench parse failed", e)
PY
