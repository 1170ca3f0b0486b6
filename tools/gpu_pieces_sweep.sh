#!/bin/bash
# Decode-step time of the bench workload for each cross-attention piece count (B200W_XT_PIECES).
for p in 1 2 3 4 6; do
  echo "pieces=$p"
  B200W_DECODE_TIMING=1 B200W_ABLATE=32768 B200W_XT_PIECES=$p python tools/ablate_decode.py --child --steps 40 2>&1 | grep "b200w_decode" | tail -1
done
