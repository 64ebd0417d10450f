#!/bin/bash
# usage: gpu_ab3d.sh EDGE V1 V2 ...: time gpurun_ab/libsmslu_<V>.so builds on the same box
EDGE=$1; shift
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do for v in "$@"; do
  SMSLU_LIB=$PWD/gpurun_ab/libsmslu_$v.so timeout 300 python scripts/ab3d.py $EDGE $v 2>&1 | grep -v Warning | tail -2
done; done | tee gpurun_out/ab3d_$(date +%H%M%S).txt
