#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?" )
tail -5 gpurun_out/r3_pytest.log
python scripts/configs_run.py lap3d 128 2>&1 | grep -v "^built"
python scripts/configs_run.py lap3d 96 2>&1 | grep -v "^built"
python scripts/configs_run.py lap2d 1024 2>&1 | grep -v "^built"
