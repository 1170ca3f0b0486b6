#!/bin/bash
# Round-2 bench check: default bench (short), small64 config, fused-step opt-in timing.
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 600 gpurun_out/r2_bench_a.err
python bench.py --config small64 --stream-seconds 30 > gpurun_out/r2_bench_small64.json 2> gpurun_out/r2_bench_small64.err; tail -c 600 gpurun_out/r2_bench_small64.err
B200W_DSTEP=1 B200W_DECODE_TIMING=1 B200W_ABLATE=32768 timeout 300 python tools/ablate_decode.py --child --steps 40 2>&1 | grep "b200w_decode" | tail -1
