#!/bin/bash
# solve-only timing with the persistent chain kernels restricted to sets of at most MAXC parallel chains
EDGE=${1:-128}
for cfg in "0 32" "1 1" "1 2" "1 4" "1 8" "1 32"; do
  set -- $cfg
  SMSLU_CHAINS=$1 SMSLU_CHAIN_MAXC=$2 timeout 300 python scripts/ab3d.py $EDGE "chains=$1,maxc=$2" 2>&1 | grep -v Warning | tail -1 | cut -c1-120
done | tee gpurun_out/chains_ab.txt
