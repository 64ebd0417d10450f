#!/bin/bash
# usage: bash scripts/gpu_profile_round.sh TAG "kre:count:skip" ...
# bench (N=1) + reference arm + ncu launch list + ncu --set full captures, all summarised ON THE BOX
# (the .ncu-rep files are deleted afterwards: gpurun_out/ is capped at 64 MiB).
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n1 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "bench ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_step.py 1024 > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
i=0
: > gpurun_out/ncu_stalls_$TAG.txt
for spec in "$@"; do
  IFS=: read KRE CNT SKIP <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$KRE -s ${SKIP:-0} -c ${CNT:-3} -o gpurun_out/prof_${TAG}_$i -f python scripts/one_step.py 1024 > gpurun_out/ncu_full_${TAG}_$i.log 2>&1
  echo "capture $i ($KRE) rc=$?"
  python scripts/ncu_summary.py gpurun_out/prof_${TAG}_$i.ncu-rep gpurun_out/ncu_full_${TAG}_$KRE.csv
  python scripts/ncu_stalls.py gpurun_out/prof_${TAG}_$i.ncu-rep >> gpurun_out/ncu_stalls_$TAG.txt
  rm -f gpurun_out/prof_${TAG}_$i.ncu-rep
  i=$((i+1))
done
du -sh gpurun_out
