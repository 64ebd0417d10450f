#!/bin/bash
# usage: gpu_run_mg.sh N : multi-GPU parity tests (world sizes up to N), then the partitioned 128^3 bench on N GPUs
N=$1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/mg_pytest_n$N.log 2>&1; echo "pytest mgpu rc=$?" )
tail -4 gpurun_out/mg_pytest_n$N.log
bash scripts/gpu_run_n.sh $N lap3d_128 3 | tail -c 1200
