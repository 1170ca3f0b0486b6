#!/bin/bash
# Decode-step time vs the cross attention's L2-prefetch distance (tiles) and piece count.
timeout 600 python -m pytest tests/test_gpu_blocks.py -m gpu -q -k "cross_attention_tc" 2>&1 | tail -2
for pf in 0 8 16 32 64; do
  echo "prefetch=$pf pieces=1"
  B200W_XT_PREFETCH=$pf B200W_DECODE_TIMING=1 B200W_ABLATE=32768 timeout 300 python tools/ablate_decode.py --child --steps 40 2>&1 | grep "b200w_decode" | tail -1
done
for p in 2 3; do
  echo "prefetch=16 pieces=$p"
  B200W_XT_PIECES=$p B200W_DECODE_TIMING=1 B200W_ABLATE=32768 timeout 300 python tools/ablate_decode.py --child --steps 40 2>&1 | grep "b200w_decode" | tail -1
done
