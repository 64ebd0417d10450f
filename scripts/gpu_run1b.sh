#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc=$?" )
tail -5 gpurun_out/r1b_pytest.log
i=0
: > gpurun_out/r1_ncu_stalls.txt
for spec in "k_gemm_cb:1:332" "k_gemm_cb:1:420" "k_fwd:2:150" "k_bwd:2:100" "k_assemble:1:119" "k_panel:2:1500"; do
  IFS=: read KRE CNT SKIP <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o gpurun_out/r1_prof_$i -f python scripts/one_step.py lap3d 128 > gpurun_out/r1_ncu_full_$i.log 2>&1
  echo "capture $i ($KRE) rc=$?"
  python scripts/ncu_summary.py gpurun_out/r1_prof_$i.ncu-rep gpurun_out/r1_ncu_full_${KRE}_$i.csv
  python scripts/ncu_stalls.py gpurun_out/r1_prof_$i.ncu-rep >> gpurun_out/r1_ncu_stalls.txt
  i=$((i+1))
done
ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
