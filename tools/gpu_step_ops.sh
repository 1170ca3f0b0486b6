#!/bin/bash
# block tests of the decode-step kernels, then step time and token checksums with the on-chip and the streaming sampler
python -m pytest tests/test_gpu_step_ops.py -x -q 2>&1 | tail -5
for s in "" 1; do
  echo "== B200W_SAMPLE_STREAMING=${s:-unset}"
  env ${s:+B200W_SAMPLE_STREAMING=1} B200W_DECODE_TIMING=1 python tools/token_checksum.py 32 5 2>&1 | grep -E "b200w_decode|tokens" | cut -c1-200 | tail -2
done
