#!/bin/bash
# ncu --set full captures of the hot kernels (one step of the bench workload, few decode steps).
# usage (on the GPU box): bash tools/ncu_full.sh <tag>
TAG=${1:-r01}
export B200W_NO_GRAPHS=1
CMD="python tools/profile_step.py --max-new 4"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:cross_attn_kernel -s 40 -c 2 -o gpurun_out/ncu_cross_attn_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:gemm_tcgen05_kernel -s 500 -c 8 -o gpurun_out/ncu_gemm_small_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:gemm_tcgen05_kernel -s 20 -c 2 -o gpurun_out/ncu_gemm_enc_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:enc_attn_kernel -s 2 -c 1 -o gpurun_out/ncu_enc_attn_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:sample_kernel -s 1 -c 1 -o gpurun_out/ncu_sample_$TAG -f $CMD > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
