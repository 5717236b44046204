#!/bin/bash
# Round-2 profiling pass (one GPU): launch list of one training step + ncu --set full captures of the main kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches.csv python scripts/profile_step.py > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
cap() {  # piece, kernel regex, skip, count, tag
  timeout 300 $NCU -k "regex:$2" -s $3 -c $4 -o gpurun_out/r02_ncu_$5 python scripts/prof_kernels.py $1 > gpurun_out/r02_ncu_$5.log 2>&1
  echo "$5 rc=$?"
}
cap fwd   gemm_kernel          3 2 fwd_gemm
cap bwd   gemm_kernel          4 2 dgrad_gemm_streamk
cap bwd   lstm_cell_bwd        4 2 cell_bwd_bias
cap rec   lstm_rec_resident    0 1 rec_resident
cap head  'gemm_kernel|moe_mix' 0 7 head
ls -la gpurun_out/r02_*.ncu-rep
