#!/bin/bash
# 8 GPUs: bench line, then a per-level profile of the partitioned run (rank 0's marks), then the multi-GPU parity tests
mkdir -p gpurun_out
bash scripts/gpu_run_n.sh 8 lap3d_128 3 | tail -c 600
( SMSLU_LEVEL_TIMES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/configs_run_mg.py lap3d 128 > gpurun_out/n8_levels.out 2> gpurun_out/n8_levels.err; echo "levels rc=$?" )
tail -2 gpurun_out/n8_levels.out
( timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/mg_pytest_n8.log 2>&1; echo "pytest mgpu rc=$?" ); tail -3 gpurun_out/mg_pytest_n8.log
