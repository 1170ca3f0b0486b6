"""Microbenchmark of the decode step's own kernels through the C ABI (GPU box):
self-attention over the paged cache (bench shape: 32 streams x 5 beams, 20 heads, rotating over 8 layers' pools so
that the cache comes from HBM as in the real step) and the fused logit rules + top-k (160 rows x 51866)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
H, K, B, NS, NL = 20, 5, 32, 4, 8
d = H * 64
rows = B * K
st = torch.cuda.current_stream().cuda_stream


def time_loop(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def attention_case(hist, shared):
    max_len = hist + 2
    per_stream = shared + K * (hist - shared + 1)
    n_slots = B * per_stream
    pools = [(torch.randn(n_slots, d, device=dev).to(torch.bfloat16), torch.randn(n_slots, d, device=dev).to(torch.bfloat16))
             for _ in range(NL)]
    table = torch.zeros(rows, max_len, dtype=torch.int32)
    slot = torch.zeros(rows, dtype=torch.int32)
    for b in range(B):
        base = b * per_stream
        for k in range(K):
            r = b * K + k
            own0 = base + shared + k * (hist - shared + 1)
            table[r, :shared] = torch.arange(base, base + shared, dtype=torch.int32)
            table[r, shared:hist] = torch.arange(own0, own0 + hist - shared, dtype=torch.int32)
            slot[r] = own0 + hist - shared
    kvlen = torch.full((rows,), hist + 1, dtype=torch.int32)
    trow = torch.arange(rows, dtype=torch.int32)
    partial = torch.randn(NS, rows, 3 * d, device=dev)
    bias = torch.randn(3 * d, device=dev)
    out = torch.zeros(rows, d, dtype=torch.bfloat16, device=dev)
    table, slot, kvlen, trow = (t.to(dev) for t in (table, slot, kvlen, trow))
    res = {}
    for name, code in (("row per warp", 1),):
        def run():
            for kp, vp in pools:
                _lib.check(lib.b200w_step_self_attention(partial.data_ptr(), NS, bias.data_ptr(), rows, d, H,
                                                         slot.data_ptr(), kp.data_ptr(), vp.data_ptr(), table.data_ptr(),
                                                         max_len, trow.data_ptr(), kvlen.data_ptr(), max_len, code,
                                                         out.data_ptr(), st))
        res[name] = time_loop(run) / NL
    unique = B * (shared + K * (hist - shared)) * d * 2 * 2 / 1e6
    print(f"self-attention hist {hist:4d} shared {shared:4d}: " + "  ".join(f"{k} {v:6.2f} us" for k, v in res.items())
          + f"   (unique K/V {unique:.1f} MB + slabs {NS * rows * 3 * d * 4 / 1e6:.1f} MB)")


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
for hist, shared in ((4, 4), (20, 17), (60, 56), (130, 126), (130, 60), (300, 296)):
    if ONLY in ("", f"attn{hist}_{shared}"):
        attention_case(hist, shared)
if ONLY not in ("", "sample"):
    sys.exit(0)

V, ld, k = 51866, 51872, 10
logits = torch.randn(rows, ld, device=dev) * 3
mask = torch.zeros(V, dtype=torch.uint8, device=dev)
state = torch.zeros(rows, 4, dtype=torch.int32, device=dev)
state[:, 0] = 1
ids = torch.zeros(rows, k, dtype=torch.int32, device=dev)
lps = torch.zeros(rows, k, device=dev)
for streaming in (0, 1):
    t = time_loop(lambda: _lib.check(lib.b200w_sample_topk(logits.data_ptr(), rows, ld, V, mask.data_ptr(), state.data_ptr(),
                                                           50257, 50363, 50365, 50, k, ids.data_ptr(), lps.data_ptr(),
                                                           streaming, 0.0, 0, st)))
    print(f"sample_topk {'streaming' if streaming else 'on-chip row'}: {t:7.2f} us ({rows} rows x {V})")
