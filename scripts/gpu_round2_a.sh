#!/bin/bash
# Round-2 evidence, call A (ONE GPU): tests, smoke, bench lines, launch list, ncu --set full captures (summarised on the box).
T=${1:-round2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
( timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" ); tail -12 gpurun_out/${T}_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" ); tail -2 gpurun_out/${T}_smoke.log
( timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?" ); head -c 600 gpurun_out/${T}_bench_n1.json; echo
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "bench ref rc=$?" ); head -c 400 gpurun_out/${T}_bench_reference.json; echo
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_one_step_128.csv python scripts/one_step.py lap3d 128 > gpurun_out/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?" )
python scripts/launch_list.py gpurun_out/${T}_launches_one_step_128.csv totals > gpurun_out/${T}_launch_totals.txt 2>&1; cat gpurun_out/${T}_launch_totals.txt
: > gpurun_out/${T}_ncu_stalls.txt
i=0
for spec in "k_gemm_cb:2:332" "k_fwd:3:30" "k_bwd:3:280" "k_assemble_smem:2:120" "k_panel:2:1500" "k_small_factor_reg:1:20"; do
  IFS=: read KRE CNT SKIP <<< "$spec"
  NAME=$KRE
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:^$KRE\$ -s $SKIP -c $CNT -o gpurun_out/${T}_prof_$i -f python scripts/one_step.py lap3d 128 > gpurun_out/${T}_ncu_full_$i.log 2>&1
  echo "capture $i ($KRE) rc=$?"
  python scripts/ncu_summary.py gpurun_out/${T}_prof_$i.ncu-rep gpurun_out/${T}_ncu_full_${NAME}_$i.csv
  python scripts/ncu_stalls.py gpurun_out/${T}_prof_$i.ncu-rep >> gpurun_out/${T}_ncu_stalls.txt
  [ $i -ne 0 ] && rm -f gpurun_out/${T}_prof_$i.ncu-rep
  i=$((i+1))
done
( timeout 400 python bench.py --config lap2d_1024 --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_lap2d_1024_n1.json 2> gpurun_out/${T}_bench_lap2d_n1.err; echo "bench lap2d rc=$?" )
( timeout 400 python scripts/rhs_sweep.py 96 2>&1 | grep -v Warn > gpurun_out/${T}_rhs_sweep_96.txt; echo "rhs sweep rc=$?" ); cat gpurun_out/${T}_rhs_sweep_96.txt
du -sh gpurun_out
