#!/bin/bash
# Final profiling pass of round 2 (one GPU): launch list of one training step + ncu --set full captures of the kernels
# that changed late in the round (cell backward with 16-byte bias reductions, weight-gradient GEMM with the sum-of-squares
# epilogue, clip+Adam with the assembled norm, classifier head with the one-round tile / split-K rules).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02d_launches.csv python scripts/profile_step.py > gpurun_out/r02d_launches.log 2>&1
echo "launch list rc=$? lines $(wc -l < gpurun_out/r02d_launches.csv)"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
cap() {
  timeout 300 $NCU -k "regex:$2" -s $3 -c $4 -o gpurun_out/r02d_ncu_$5 python scripts/prof_kernels.py $1 > gpurun_out/r02d_ncu_$5.log 2>&1
  echo "$5 rc=$?"
}
cap bwd   lstm_cell_bwd        4 2 cell_bwd
cap wgrad gemm_kernel          0 2 wgrad_sumsq
cap adam  clip_adam            0 3 clip_adam
cap head  'gemm_kernel|moe_mix|reg_cross' 0 9 head
ls -la gpurun_out/r02d_*.ncu-rep
