#!/bin/bash
# Fused decode step (dstep.cu) vs the launch-per-op step: token checksums, then the step time of both.
set -x
timeout 300 python tools/token_checksum.py 8 5
B200W_NO_DSTEP=1 timeout 300 python tools/token_checksum.py 8 5
for v in 0 1; do
  if [ $v = 1 ]; then export B200W_NO_DSTEP=1; else unset B200W_NO_DSTEP; fi
  B200W_DECODE_TIMING=1 B200W_ABLATE=32768 timeout 300 python tools/ablate_decode.py --child --steps 40 2>&1 | grep "b200w_decode" | tail -1
done
unset B200W_NO_DSTEP
timeout 900 python -m pytest tests/test_gpu_parity_sweep.py tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -i "chunks, \|DTW path identical\|passed\|failed\|error" | cut -c1-250
