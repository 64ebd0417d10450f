#!/bin/bash
# time several builds of libsmslu.so on the same box: gpurun_ab/libsmslu_<V>.so for V in "$@"
for rep in 1 2; do for v in "$@"; do
SMSLU_LIB=$PWD/gpurun_ab/libsmslu_$v.so python scripts/split_time.py 2>&1 | tr '\n' ' ' | sed "s/^/$v rep $rep: /"; echo
done; done
