#!/bin/bash
# chain solve kernels: correctness on small/medium cases under a timeout, then timing
mkdir -p gpurun_out
( SMSLU_CHAINS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_properties or north_star or many_rhs or factor_and_solve" > gpurun_out/r6_pytest.log 2>&1; echo "pytest rc=$?" )
tail -15 gpurun_out/r6_pytest.log
SMSLU_CHAINS=1 timeout 120 python scripts/split_time3d.py 96 2>&1 | grep -v "^built"
SMSLU_CHAINS=1 timeout 200 python scripts/split_time3d.py 128 2>&1 | grep -v "^built"
timeout 200 python scripts/split_time3d.py 128 2>&1 | grep -v "^built"
SMSLU_CHAINS=1 timeout 120 python scripts/split_time.py 1024 2>&1 | grep -v "^built"
SMSLU_GEMM_STRIP=8 timeout 200 python scripts/split_time3d.py 128 2>&1 | grep -v "^built"
SMSLU_GEMM_STRIP=16 timeout 200 python scripts/split_time3d.py 128 2>&1 | grep -v "^built"
SMSLU_GEMM_STRIP=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "factor_and_solve or refactor_with_new" 2>&1 | tail -3
