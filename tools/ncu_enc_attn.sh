#!/bin/bash
# ncu --set full of the encoder attention kernel on tools/enc_attn_bench.py (variant from the environment)
TAG=${1:-r02b}
python tools/enc_attn_bench.py > gpurun_out/enc_attn_bench_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/enc_attn_bench_$TAG.log; exit 1; }
cat gpurun_out/enc_attn_bench_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:enc_attn -s 4 -c 1 -o gpurun_out/ncu_enc_attn2_$TAG -f python tools/enc_attn_bench.py > /dev/null 2>&1
ls -la gpurun_out/ncu_enc_attn2_$TAG.ncu-rep
