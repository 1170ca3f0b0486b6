#!/bin/bash
# block tests of the decode-step kernels, then step time and token checksums with 1 / K rows per self-attention CTA
python -m pytest tests/test_gpu_step_ops.py -x -q 2>&1 | tail -15
for w in 1 5; do
  echo "== B200W_QKV_ROWS=$w"
  B200W_QKV_ROWS=$w B200W_DECODE_TIMING=1 python tools/token_checksum.py 32 5 2>&1 | grep -E "b200w_decode|tokens" | cut -c1-200 | tail -3
done
echo "== streaming sample kernel"
B200W_SAMPLE_STREAMING=1 B200W_DECODE_TIMING=1 python tools/token_checksum.py 32 5 2>&1 | grep -E "b200w_decode|tokens" | cut -c1-200 | tail -3
