#!/bin/bash
# ncu capture of the Schur GEMM after the fast-path work + solve-overhead experiments
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_cb -s 332 -c 1 -o gpurun_out/r4_prof_gemm -f python scripts/one_step.py lap3d 128 > gpurun_out/r4_ncu_gemm.log 2>&1
echo "capture rc=$?"
python scripts/ncu_summary.py gpurun_out/r4_prof_gemm.ncu-rep gpurun_out/r4_ncu_full_k_gemm_cb.csv
python scripts/ncu_stalls.py gpurun_out/r4_prof_gemm.ncu-rep
for v in "" "SMSLU_NO_LANES=1" "SMSLU_NO_PDL=1"; do
  echo "== solve timing with [$v]"
  env $v python scripts/split_time3d.py 128 2>&1 | grep -v "^built"
done
