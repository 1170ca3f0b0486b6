#!/bin/bash
# self-attention of the decode step: rows (beams) of one stream per CTA, 1 vs K; step time and token checksums
for w in 1 5; do
  echo "== B200W_QKV_ROWS=$w"
  B200W_QKV_ROWS=$w B200W_DECODE_TIMING=1 python tools/token_checksum.py 32 5 2>&1 | grep -E "b200w_decode|tokens" | tail -3
done
