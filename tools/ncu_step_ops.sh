#!/bin/bash
# ncu --set full of the decode step's self-attention and sampling kernels on the microbenchmark (tools/step_ops_bench.py)
TAG=${1:-r02b}
python tools/step_ops_bench.py > gpurun_out/step_ops_bench_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/step_ops_bench_$TAG.log; exit 1; }
cat gpurun_out/step_ops_bench_$TAG.log
NCU="ncu --set full --clock-control none --import-source on"
# (each case runs 184 launches: 3 warm-ups + 20 repetitions over 8 layers' pools)
$NCU -k regex:qkv_attn_kernel -s 60 -c 1 -o gpurun_out/ncu_qkv_attn_h130_$TAG -f python tools/step_ops_bench.py attn130_126 > /dev/null 2>&1
$NCU -k regex:qkv_attn_kernel -s 60 -c 1 -o gpurun_out/ncu_qkv_attn_h4_$TAG -f python tools/step_ops_bench.py attn4_4 > /dev/null 2>&1
$NCU -k regex:sample_row_kernel -s 8 -c 1 -o gpurun_out/ncu_sample_row_$TAG -f python tools/step_ops_bench.py sample > /dev/null 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
