#!/bin/bash
# 2 GPUs: CUDA IPC probe, then the partitioned CUDA path (distributed top) against the oracle
mkdir -p gpurun_out
nvidia-smi -L
( timeout 120 scripts/micro/ipc_probe > gpurun_out/r2_ipc_probe.txt 2>&1; echo "ipc probe rc=$?" )
cat gpurun_out/r2_ipc_probe.txt
( timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$?" )
tail -30 gpurun_out/r2_pytest_mgpu.log
