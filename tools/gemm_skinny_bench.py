"""Decode-shaped (skinny) GEMM micro-benchmark: M rows x the four weight shapes of a decoder layer,
cycling over 32 distinct weight matrices (HBM-cold, like a decode step), timed with CUDA events over a
CUDA-graph replay of the 32 launches.  Optionally dumps the per-CTA timestamps (b200w_gemm_set_trace).
usage (GPU box): python tools/gemm_skinny_bench.py [--m 160] [--trace]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=160)
ap.add_argument("--trace", action="store_true")
ap.add_argument("--skinny", action="store_true")
ap.add_argument("--floor", action="store_true")
ap.add_argument("--sktrace", action="store_true")
a = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
M, L = a.m, 32
SHAPES = {"qkv": (3840, 1280), "out": (1280, 1280), "fc1": (5120, 1280), "fc2": (1280, 5120)}
CONFIGS = [  # (force_bn, splits)
    (64, 1), (64, 2), (64, 3), (64, 4), (64, 5), (64, 7), (64, 8), (64, 10), (64, 16), (128, 1), (128, 4), (128, 8)]


def run(N, K, bn, sp, trace=None):
    W = torch.randn(L, N, K, device=dev, dtype=torch.bfloat16) * 0.02
    A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(max(sp, 1) * M * N, device=dev, dtype=torch.float32)
    flags = 1 | ((sp if sp > 1 else 0) << 8)
    s = torch.cuda.Stream()
    def enqueue():
        for l in range(L):
            _lib.check(lib.b200w_gemm_bf16(A.data_ptr(), W[l].data_ptr(), None, None, out.data_ptr(), M, N, K, K, K,
                                           0, N, flags, bn, s.cuda_stream))
    with torch.cuda.stream(s):
        enqueue()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            enqueue()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (5 * L)
    # correctness of the last launch
    ref = A.float() @ W[L - 1].float().t()
    got = out.view(max(sp, 1), M, N).sum(0)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    return us, err


if a.trace:
    tb = torch.zeros(148 * 16, device=dev, dtype=torch.int64)
    for name, (N, K) in SHAPES.items():
        for bn, sp in [(64, 1), (64, 5)]:
            W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02
            A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            out = torch.empty(max(sp, 1) * M * N, device=dev, dtype=torch.float32)
            flags = 1 | ((sp if sp > 1 else 0) << 8)
            torch.cuda.synchronize()
            for it in range(3):
                tb.zero_()
                torch.empty(64 << 20, device=dev, dtype=torch.float32).fill_(1.0)  # flush L2
                torch.cuda.synchronize()
                lib.b200w_gemm_set_trace(tb.data_ptr())
                _lib.check(lib.b200w_gemm_bf16(A.data_ptr(), W.data_ptr(), None, None, out.data_ptr(), M, N, K, K, K, 0,
                                               N, flags, bn, None))
                torch.cuda.synchronize()
                lib.b200w_gemm_set_trace(None)
            t = tb.cpu().numpy().reshape(148, 16)
            t = t[t[:, 1] != 0]
            c = t[:, 1:9] - t[:, 1:2]
            wall = (t[:, 9].max() - t[:, 0].min()) / 1000.0
            start_spread = (t[:, 0].max() - t[:, 0].min()) / 1000.0
            med = np.median(c, axis=0)
            print(f"{name:4s} bn={bn} sp={sp} ctas={len(t):3d} wall {wall:6.2f} us (CTA start spread {start_spread:5.2f} us) "
                  f"cycles: prologue {med[1]:6.0f} | wait {med[2]:6.0f} | first full {med[3]:6.0f} | mma issued {med[4]:6.0f} | "
                  f"tfull {med[5]:6.0f} | epi done {med[6]:6.0f} | exit {med[7]:6.0f}")
    sys.exit(0)

def run_skinny(N, K, sp, t64=False):
    W = torch.randn(L, N, K, device=dev, dtype=torch.bfloat16) * 0.02
    A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(max(sp, 1) * M * N, device=dev, dtype=torch.float32)
    s = torch.cuda.Stream()
    def enqueue():
        for l in range(L):
            _lib.check(lib.b200w_skinny_gemm(A.data_ptr(), W[l].data_ptr(), None, out.data_ptr(), M, N, K, K, K, N, 1,
                                             sp | ((64 << 8) if t64 else 0), s.cuda_stream))
    with torch.cuda.stream(s):
        enqueue()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            enqueue()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (5 * L)
    ref = A.float() @ W[L - 1].float().t()
    got = out.view(max(sp, 1), M, N).sum(0)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    return us, err


if "--sktrace" in sys.argv:  # run with B200W_TRACE_SKINNY=1
    tb = torch.zeros(160 * 16, device=dev, dtype=torch.int64)
    for (N, K, sp) in [(128, 128, 1), (1280, 1280, 10), (3840, 1280, 4), (1280, 5120, 14)]:
        W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(max(sp, 1) * M * N, device=dev, dtype=torch.float32)
        for it in range(3):
            tb.zero_()
            torch.empty(64 << 20, device=dev, dtype=torch.float32).fill_(1.0)
            torch.cuda.synchronize()
            lib.b200w_gemm_set_trace(tb.data_ptr())
            _lib.check(lib.b200w_skinny_gemm(A.data_ptr(), W.data_ptr(), None, out.data_ptr(), M, N, K, K, K, N, 1, sp, None))
            torch.cuda.synchronize()
            lib.b200w_gemm_set_trace(None)
        t = tb.cpu().numpy().reshape(160, 16)
        t = t[t[:, 1] != 0]
        c = np.median(t[:, 2:9] - t[:, 1:2], axis=0)
        wall = (t[:, 9].max() - t[:, 0].min()) / 1000.0
        print(f"N={N} K={K} sp={sp} ctas={len(t)} wall(first start..last end) {wall:6.2f} us | cycles from CTA start: params+alloc {c[0]:6.0f} | "
              f"producer past wait {c[1]:6.0f} | first full {c[2]:6.0f} | mma committed {c[3]:6.0f} | epi start {c[4]:6.0f} | epi stores issued {c[5]:6.0f} | exit {c[6]:6.0f}"
              f" || chunk0 ld {np.median(t[:, 10] - t[:, 1]):6.0f} st {np.median(t[:, 11] - t[:, 1]):6.0f} chunk1 ld {np.median(t[:, 12] - t[:, 1]):6.0f} st {np.median(t[:, 13] - t[:, 1]):6.0f} chunk2 ld {np.median(t[:, 14] - t[:, 1]):6.0f} st {np.median(t[:, 15] - t[:, 1]):6.0f}")
    sys.exit(0)

if "--floor" in sys.argv:
    for (N, K, sp) in [(128, 128, 1), (1280, 128, 1), (1280, 256, 1), (1280, 1280, 1), (12800, 128, 1), (18944, 128, 1), (18944, 256, 1), (18944, 640, 1)]:
        us, err = run_skinny(N, K, sp)
        print(f"floor skinny N={N:5d} K={K:5d} splits={sp} ctas={(N // 128) * sp:3d}: {us:7.2f} us/launch relerr {err:.1e}", flush=True)
    for (N, K, bn, sp) in [(128, 128, 64, 1), (1280, 128, 64, 1), (1280, 1280, 64, 5), (9472, 128, 64, 1)]:
        us, err = run(N, K, bn, sp)
        print(f"floor gemm   N={N:5d} K={K:5d} bn={bn} splits={sp}: {us:7.2f} us/launch relerr {err:.1e}", flush=True)
    # empty-ish torch kernels in the same graph pattern for reference
    x = torch.zeros(1024, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(L):
                x.add_(1.0)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s); s.synchronize()
    print(f"floor torch add_ (1 CTA): {e0.elapsed_time(e1) * 1000 / (5 * L):7.2f} us/launch")
    sys.exit(0)

if "--skinny" in sys.argv:
    for name, (N, K) in SHAPES.items():
        for sp in (1, 2, 3, 4, 5, 7, 8, 10, 14):
            if K // 64 < sp * 2 or (N // 128) * sp > 160:
                continue
            us, err = run_skinny(N, K, sp)
            print(f"skinny128 {name:4s} N={N:5d} K={K:5d} splits={sp:2d} ctas={(N // 128) * sp:3d}: {us:7.2f} us/launch  {N * K * 2 / us / 1e3:7.1f} GB/s weights  relerr {err:.1e}", flush=True)
        for sp in (1, 2, 3, 4, 5, 7):
            if K // 64 < sp * 2 or (N // 64) * sp > 160:
                continue
            us, err = run_skinny(N, K, sp, True)
            print(f"skinny64  {name:4s} N={N:5d} K={K:5d} splits={sp:2d} ctas={(N // 64) * sp:3d}: {us:7.2f} us/launch  {N * K * 2 / us / 1e3:7.1f} GB/s weights  relerr {err:.1e}", flush=True)
    sys.exit(0)

for name, (N, K) in SHAPES.items():
    mb = N * K * 2 / 1e6
    for bn, sp in CONFIGS:
        if K // 64 < sp * 2:
            continue
        us, err = run(N, K, bn, sp)
        print(f"{name:4s} N={N:5d} K={K:5d} bn={bn:3d} splits={sp:2d}: {us:7.2f} us/launch  {mb / us * 1e3 / 1e3:7.1f} GB/s weights  relerr {err:.1e}",
              flush=True)
