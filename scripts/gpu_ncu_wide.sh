#!/bin/bash
# ncu --set full captures of the 32-wide solve kernels on bulk levels (96^3)
T=${1:-round2_final}
mkdir -p gpurun_out
: > gpurun_out/${T}_ncu_stalls_wide.txt
i=0
for spec in "k_fwd32:2:12" "k_bwd32:2:130"; do
  IFS=: read KRE CNT SKIP <<< "$spec"
  timeout 100 ncu --set full --clock-control none --import-source on -k regex:^$KRE\$ -s $SKIP -c $CNT -o gpurun_out/${T}_wide_prof_$i -f python scripts/one_step_wide.py 96 > gpurun_out/${T}_ncu_wide_$i.log 2>&1
  echo "capture $i ($KRE) rc=$?"
  python scripts/ncu_summary.py gpurun_out/${T}_wide_prof_$i.ncu-rep gpurun_out/${T}_ncu_full_${KRE}.csv
  python scripts/ncu_stalls.py gpurun_out/${T}_wide_prof_$i.ncu-rep >> gpurun_out/${T}_ncu_stalls_wide.txt
  rm -f gpurun_out/${T}_wide_prof_$i.ncu-rep
  i=$((i+1))
done
cat gpurun_out/${T}_ncu_stalls_wide.txt
