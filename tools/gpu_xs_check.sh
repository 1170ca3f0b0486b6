#!/bin/bash
# Single-pass cross attention: block test, checksums vs the two-pass kernel, decode-step time for piece counts.
timeout 600 python -m pytest tests/test_gpu_blocks.py -m gpu -q -k "cross_attention_tc" 2>&1 | tail -3
timeout 300 python tools/token_checksum.py 8 5 2>&1 | tail -1
B200W_XT_TWO_PASS=1 timeout 300 python tools/token_checksum.py 8 5 2>&1 | tail -1
for p in 1 2 3; do
  echo "single-pass pieces=$p"
  B200W_DECODE_TIMING=1 B200W_ABLATE=32768 B200W_XT_PIECES=$p timeout 300 python tools/ablate_decode.py --child --steps 40 2>&1 | grep "b200w_decode" | tail -1
done
echo "two-pass pieces=1"
B200W_XT_TWO_PASS=1 B200W_DECODE_TIMING=1 B200W_ABLATE=32768 timeout 300 python tools/ablate_decode.py --child --steps 40 2>&1 | grep "b200w_decode" | tail -1
