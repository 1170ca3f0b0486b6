#!/bin/bash
python -m pytest tests/test_gpu_blocks.py -x -q -k "enc_attention or cross_attention_flash" 2>&1 | tail -3
B200W_ENC_ATTN_V1=1 python tools/enc_attn_bench.py
for pe in ${PES:-0 8}; do B200W_ENC_ATTN_PE=$pe timeout 120 python tools/enc_attn_bench.py; done
