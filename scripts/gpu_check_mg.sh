#!/bin/bash
# usage (on a multi-GPU box): bash scripts/gpu_check_mg.sh TAG NGPU
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "bench n$N rc=$?"; tail -5 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_${TAG}_n%d.json" % n) if l.startswith("{")][-1])
        print("N=%d ms/step %.3f e2e %.3f res %.2e %s" % (n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["residual"], d["scaling"]))
        print({k: v["ms"] for k, v in d["kernels"].items()})
    except Exception as e:
        print("bench parse failed", n, e)
PY
