#!/bin/bash
# usage: bash scripts/gpu_prof2.sh TAG "kre1:count1:skip1" "kre2:count2:skip2" ...   (ncu --set full captures of one_step.py)
TAG=$1; shift
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  IFS=: read KRE CNT SKIP <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$KRE -s ${SKIP:-0} -c ${CNT:-3} -o gpurun_out/prof_${TAG}_$i -f python scripts/one_step.py 1024 > gpurun_out/ncu_full_${TAG}_$i.log 2>&1
  echo "capture $i ($KRE) rc=$?"
  i=$((i+1))
done
