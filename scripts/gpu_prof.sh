#!/bin/bash
# usage: bash scripts/gpu_prof.sh TAG [kernel-regex for the --set full capture] [max launches]
TAG=${1:-x}; KRE=${2:-}; CNT=${3:-12}
mkdir -p gpurun_out
python scripts/one_step.py 1024 > gpurun_out/one_step_$TAG.log 2>&1 || { echo one_step failed; tail -5 gpurun_out/one_step_$TAG.log; exit 1; }
tail -1 gpurun_out/one_step_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_step.py 1024 > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
if [ -n "$KRE" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$KRE -c $CNT -o gpurun_out/prof_${TAG} -f python scripts/one_step.py 1024 > gpurun_out/ncu_full_$TAG.log 2>&1
  echo "full capture rc=$?"
fi
