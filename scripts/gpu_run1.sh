#!/bin/bash
# round-2 GPU run 1 (one GPU): tests, 128^3 bench line, level profile, launch list, ncu captures at 128^3
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" )
tail -5 gpurun_out/r1_pytest.log
( timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err; echo "bench rc=$?" )
tail -c 1500 gpurun_out/r1_bench_n1.json
( timeout 600 python scripts/level_times.py lap3d 128 > gpurun_out/r1_levels.out 2> gpurun_out/r1_levels.err; echo "levels rc=$?" )
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1_launches_128.csv python scripts/one_step.py lap3d 128 > gpurun_out/r1_ncu_launches.log 2>&1; echo "launch list rc=$?" )
python scripts/launch_list.py gpurun_out/r1_launches_128.csv > gpurun_out/r1_launch_totals.txt 2>&1
du -sh gpurun_out
