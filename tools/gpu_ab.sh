#!/bin/bash
# A/B on one box: decode step time with the library in lib/ and with lib/libb200whisper_ab.so (B200W_LIB)
for rep in 1 2; do
for lib in "" realtime-whisper-asr_b200/lib/libb200whisper_ab.so; do
  echo "== lib=${lib:-default}"
  B200W_LIB=$lib B200W_DECODE_TIMING=1 python tools/token_checksum.py 32 5 2>&1 | grep -E "b200w_decode|tokens" | cut -c1-160 | tail -2
done
done
