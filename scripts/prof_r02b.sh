#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches.csv python scripts/profile_step.py > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$? lines $(wc -l < gpurun_out/r02_launches.csv)"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
cap() {
  timeout 300 $NCU -k "regex:$2" -s $3 -c $4 -o gpurun_out/r02_ncu_$5 python scripts/prof_kernels.py $1 > gpurun_out/r02_ncu_$5.log 2>&1
  echo "$5 rc=$?"
}
cap fwd   gemm_kernel          3 2 fwd_gemm
cap bwd   gemm_kernel          4 2 dgrad_gemm_streamk
