#!/bin/bash
# Round-2 final records: GPU suite, default bench, small64 config, parity sweep summary, short launch list
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; tail -3 gpurun_out/r2_final_tests.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 300 gpurun_out/r2_final_bench.err
python bench.py --config small64 --stream-seconds 30 > gpurun_out/r2_final_small64.json 2> gpurun_out/r2_final_small64.err; tail -c 200 gpurun_out/r2_final_small64.err
python -m pytest tests/test_gpu_parity_sweep.py -q -s 2>&1 | grep -E "chunks,|DTW path identical|passed|failed|diverge" > gpurun_out/r2_final_sweep.txt; cat gpurun_out/r2_final_sweep.txt
export B200W_NO_GRAPHS=1
CMD="python tools/profile_step.py --max-new 6"
$CMD > gpurun_out/ncu_plain_r02g.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02g.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:qkv_attn_kernel -s 40 -c 1 -o gpurun_out/ncu_qkv_attn_r02g -f $CMD > /dev/null 2>&1
ls -la gpurun_out/launches_r02g.csv gpurun_out/ncu_qkv_attn_r02g.ncu-rep
