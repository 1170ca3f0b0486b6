#!/bin/bash
# ncu captures of the hot kernels over one step of the bench workload with a few decode steps.
# usage (on the GPU box): bash tools/ncu_full.sh <tag>     (run the plain command first: it must exit 0)
TAG=${1:-r02}
export B200W_NO_GRAPHS=1
CMD="python tools/profile_step.py --max-new 6"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.log; exit 1; }
# launch list (durations of every kernel of the step, serialised, cold caches)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > /dev/null 2>&1
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
$NCU -k regex:cross_attn_sp_kernel -s 40 -c 2 -o gpurun_out/ncu_cross_sp_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:skinny_gemm_kernel -s 200 -c 6 -o gpurun_out/ncu_skinny_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:gemm_tcgen05_kernel -s 20 -c 4 -o gpurun_out/ncu_gemm_enc_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:enc_attn -s 2 -c 1 -o gpurun_out/ncu_enc_attn_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:qkv_attn_kernel -s 40 -c 1 -o gpurun_out/ncu_qkv_attn_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:sample_row_kernel -s 2 -c 1 -o gpurun_out/ncu_sample_row_$TAG -f $CMD > /dev/null 2>&1
$NCU -k regex:logmel_kernel -c 1 -o gpurun_out/ncu_logmel_$TAG -f $CMD > /dev/null 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
