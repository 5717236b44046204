#!/bin/bash
# ncu --set full captures of the step's main kernels (one process per kernel family; see prof_kernels.py)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
cap() {  # piece, kernel regex, skip, count, tag
  timeout 240 $NCU -k "regex:$2" -s $3 -c $4 -o gpurun_out/r01_ncu_$5 python scripts/prof_kernels.py $1 > gpurun_out/r01_ncu_$5.log 2>&1
  echo "$5 rc=$?"
}
cap fwd   gemm_kernel          3 2 fwd_gemm
cap bwd   gemm_kernel          4 2 dgrad_gemm
cap bwd   lstm_cell_bwd        4 2 cell_bwd
cap wgrad gemm_kernel          0 3 wgrad_dx_gemm
cap wgrad colsum               0 1 colsum
cap adam  'clip_adam|sumsq'    8 4 adam
cap head  'gemm_kernel|moe_mix' 0 7 head
ls -la gpurun_out/*.ncu-rep
