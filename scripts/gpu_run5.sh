#!/bin/bash
mkdir -p gpurun_out
python scripts/split_time3d.py 128 2>&1 | grep -v "^built"
python scripts/split_time3d.py 96 2>&1 | grep -v "^built"
( timeout 600 python scripts/level_times.py lap3d 128 > gpurun_out/r5_levels.out 2> gpurun_out/r5_levels.err; echo "levels rc=$?" )
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r5_launches_128.csv python scripts/one_step.py lap3d 128 > gpurun_out/r5_ncu_launches.log 2>&1; echo "launch list rc=$?" )
