#!/bin/bash
mkdir -p gpurun_out
( SMSLU_CHAINS=1 timeout 300 python scripts/level_times.py lap3d 128 > gpurun_out/r8_levels.out 2> gpurun_out/r8_levels.err; echo "levels rc=$?" )
