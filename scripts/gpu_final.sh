#!/bin/bash
# final validation of a round: GPU tests, smoke, default bench (N=1) and the reference arm, bench at N GPUs
TAG=${1:-final}; N=${2:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n1 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "bench ref rc=$?"
if [ "$N" -gt 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench n$N rc=$?"
fi
python - <<PY
import json
for n in ("n1", "ref", "n$N"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_${TAG}_%s.json" % n) if l.startswith("{")][-1])
        print(n, "value %.3f %s ms/step %.3f e2e %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"]), d.get("roofline", {}).get("kernel"), d.get("roofline", {}).get("frac"), d.get("clocks"))
        if "cpu_baseline" in d: print("   cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:120])
    except Exception as e:
        print("parse failed", n, e)
PY
