#!/bin/bash
mkdir -p gpurun_out
( SMSLU_CHAINS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_properties or north_star or factor_and_solve" > gpurun_out/r7_pytest.log 2>&1; echo "pytest rc=$?" )
tail -3 gpurun_out/r7_pytest.log
SMSLU_CHAINS=1 timeout 120 python scripts/split_time3d.py 96 2>&1 | grep -v "^built"
SMSLU_CHAINS=1 timeout 200 python scripts/split_time3d.py 128 2>&1 | grep -v "^built"
