#!/bin/bash
# ncu --set full captures of the bulk-level solve launches (one step of 128^3)
T=${1:-round2b}
mkdir -p gpurun_out
: > gpurun_out/${T}_ncu_stalls_solve.txt
i=0
for spec in "k_fwd:3:30" "k_bwd:3:280"; do
  IFS=: read KRE CNT SKIP <<< "$spec"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:^$KRE\$ -s $SKIP -c $CNT -o gpurun_out/${T}_solve_prof_$i -f python scripts/one_step.py lap3d 128 > gpurun_out/${T}_ncu_solve_$i.log 2>&1
  echo "capture $i ($KRE) rc=$?"
  python scripts/ncu_summary.py gpurun_out/${T}_solve_prof_$i.ncu-rep gpurun_out/${T}_ncu_full_${KRE}.csv
  python scripts/ncu_stalls.py gpurun_out/${T}_solve_prof_$i.ncu-rep >> gpurun_out/${T}_ncu_stalls_solve.txt
  i=$((i+1))
done
cat gpurun_out/${T}_ncu_stalls_solve.txt
