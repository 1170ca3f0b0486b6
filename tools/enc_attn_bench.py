"""Encoder self-attention at the bench shape (32 windows x 1500 frames x 20 heads) through the C ABI: time per launch
and TFLOP/s.  Kernel variant from the environment (B200W_ENC_ATTN_V1=1, B200W_ENC_ATTN_PE=0|2|4|8)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
B, T, H = 32, 1500, 20
d = H * 64
qkv = (torch.randn(B, T, 3 * d, device=dev) * 0.8).to(torch.bfloat16)
out = torch.zeros(B, T, d, dtype=torch.bfloat16, device=dev)
st = torch.cuda.current_stream().cuda_stream
run = lambda: _lib.check(lib.b200w_enc_attention(qkv.data_ptr(), B, T, d, H, out.data_ptr(), st))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
flops = 4.0 * T * T * d * B
# accuracy vs torch fp32 on one window / 4 heads
q, k, v = (qkv[0, :, i * d:(i + 1) * d].float().view(T, H, 64)[:, :4] for i in range(3))
ref = torch.einsum("hqk,khc->qhc", torch.softmax(torch.einsum("qhc,khc->hqk", q, k) * 0.125, -1), v)
err = (out[0].float().view(T, H, 64)[:, :4] - ref).abs().max().item()
print(f"enc attention {os.environ.get('B200W_ENC_ATTN_V1') and 'v1' or 'v2 PE=' + os.environ.get('B200W_ENC_ATTN_PE', '8')}: "
      f"{us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  max abs err {err:.4f}")

if int(os.environ.get("B200W_ENC_ATTN_DBG", "0")) & 2:
    import numpy as np
    tr = np.zeros(4 * 32 * 8, dtype=np.int64)
    _lib.check(lib.b200w_enc_attention_trace_read(tr.ctypes.data, tr.size))
    tr = tr.reshape(4, 32, 8)
    t0 = tr[2, 0, 0]
    rel = lambda x: int(x - t0) if x else -1
    print("cycles relative to the MMA issuer's first stamp (one CTA)")
    print("blk | softmax0: waitS gotS ld'd math'd gotO arrived | softmax1: ... | mma: top gotK issuedS gotV gotP0 gotP1 pvdone | tma: top gotKfree gotVfree")
    for j in range(24):
        row = [f"{j:3d} |"]
        for role, ns in ((0, 6), (1, 6), (2, 7), (3, 3)):
            row.append(" ".join(f"{rel(tr[role, j, k]):7d}" for k in range(ns)) + " |")
        print(" ".join(row))
