#!/bin/bash
# usage (on the GPU box): bash scripts/gpu_check.sh TAG   -> gpurun_out/{pytest,bench,sanitizer}_TAG.*
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python scripts/one_step.py 64 > gpurun_out/sanitizer_$TAG.log 2>&1
echo "sanitizer rc=$?" 
tail -3 gpurun_out/sanitizer_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$TAG.json"))
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "res", d["residual"])
    for k,v in d["kernels"].items(): print(k, v)
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
