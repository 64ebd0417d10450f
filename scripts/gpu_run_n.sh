#!/bin/bash
# usage: gpu_run_n.sh N [config] [steps]: partitioned bench line on N GPUs (+ the multi-GPU parity tests when N == 2)
N=$1; CFG=${2:-lap3d_128}; STEPS=${3:-3}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $STEPS --warmup 3 --no-cpu --config $CFG > gpurun_out/rn_bench_${CFG}_n$N.json 2> gpurun_out/rn_bench_${CFG}_n$N.err
echo "bench N=$N rc=$?"
tail -c 2500 gpurun_out/rn_bench_${CFG}_n$N.json
grep -v "^built\|Warning\|warn" gpurun_out/rn_bench_${CFG}_n$N.err | tail -15
