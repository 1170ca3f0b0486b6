"""GPU parity tests of the building blocks, through the C ABI (ctypes), vs plain torch fp32."""
import numpy as np
import pytest
import torch

from realtime_whisper_asr_b200 import _lib

pytestmark = pytest.mark.gpu


def _gemm(lib, A, W, bias=None, res=None, out_fp32=False, gelu=False, force_bn=0):
    M, K = A.shape
    N = W.shape[0]
    ldo = (N + 7) // 8 * 8
    out = torch.zeros((M, ldo), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=A.device)
    flags = (1 if out_fp32 else 0) | (2 if (res is not None and res.dtype == torch.float32) else 0) | (4 if gelu else 0)
    _lib.check(lib.b200w_gemm_bf16(A.data_ptr(), W.data_ptr(),
                                   bias.data_ptr() if bias is not None else None,
                                   res.data_ptr() if res is not None else None, out.data_ptr(),
                                   M, N, K, A.stride(0), W.stride(0),
                                   res.stride(0) if res is not None else 0, ldo, flags, force_bn,
                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out[:, :N]


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (128, 256, 128, 256), (128, 64, 64, 64),
    (300, 384, 384, 0), (1500, 1280, 1280, 0), (1500, 3840, 1280, 256), (777, 1000, 512, 128),
    (5, 1280, 1280, 0), (160, 5120, 1280, 64), (33, 51866, 384, 0), (3000, 384, 1536, 0),
])
def test_gemm_plain(lib, cuda_device, M, N, K, bn):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda_device)
    out = _gemm(lib, A, W, out_fp32=True, force_bn=bn)
    ref = A.float() @ W.float().T
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_gemm_epilogues(lib, cuda_device):
    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 515, 768, 256
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    res32 = torch.randn(M, N, generator=g).to(cuda_device)
    res16 = res32.to(torch.bfloat16)
    base = A.float() @ W.float().T + bias
    # bias + GELU -> bf16
    out = _gemm(lib, A, W, bias=bias, gelu=True).float()
    ref = torch.nn.functional.gelu(base)
    assert (out - ref).abs().max().item() <= 2e-2
    # bias + fp32 residual -> fp32
    out = _gemm(lib, A, W, bias=bias, res=res32, out_fp32=True)
    assert (out - (base + res32)).abs().max().item() <= 2e-3 * base.abs().max().item()
    # bf16 residual -> bf16
    out = _gemm(lib, A, W, bias=bias, res=res16).float()
    assert (out - (base + res16.float())).abs().max().item() <= 5e-2


@pytest.mark.parametrize("stride,B,T,C,N", [(1, 2, 3000, 128, 384), (2, 2, 3000, 384, 384), (1, 1, 200, 64, 128), (2, 3, 256, 128, 256)])
def test_conv1d(lib, cuda_device, stride, B, T, C, N):
    g = torch.Generator(device="cpu").manual_seed(11 + stride)
    x = (torch.randn(B, T, C, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = (torch.randn(N, C, 3, generator=g) * 0.05).to(torch.bfloat16).to(cuda_device)  # torch layout
    bias = torch.randn(N, generator=g).to(cuda_device)
    T_out = T // stride
    addend = torch.randn(T_out, N, generator=g).to(cuda_device)
    w_tap = w.permute(0, 2, 1).contiguous().view(N, 3 * C)
    out = torch.zeros((B, T_out, N), dtype=torch.float32, device=cuda_device)
    _lib.check(lib.b200w_conv1d_k3(x.data_ptr(), w_tap.data_ptr(), bias.data_ptr(), addend.data_ptr(),
                                   out.data_ptr(), B, T, C, N, stride, 1,
                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias, stride=stride, padding=1)
    ref = torch.nn.functional.gelu(ref).transpose(1, 2) + addend
    err = (out - ref).abs().max().item()
    assert err <= 5e-3, f"max abs err {err}"


@pytest.mark.parametrize("B,T,H", [(1, 1500, 6), (2, 1500, 2), (1, 128, 1), (2, 333, 4)])
def test_enc_attention(lib, cuda_device, B, T, H):
    d = H * 64
    g = torch.Generator(device="cpu").manual_seed(B * 100 + T)
    qkv = (torch.randn(B, T, 3 * d, generator=g)).to(torch.bfloat16).to(cuda_device)
    out = torch.zeros((B, T, d), dtype=torch.bfloat16, device=cuda_device)
    _lib.check(lib.b200w_enc_attention(qkv.data_ptr(), B, T, d, H, out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    q, k, v = qkv.float().split(d, dim=-1)
    q = q.view(B, T, H, 64).transpose(1, 2)
    k = k.view(B, T, H, 64).transpose(1, 2)
    v = v.view(B, T, H, 64).transpose(1, 2)
    ref = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v
    ref = ref.transpose(1, 2).reshape(B, T, d)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2, f"max abs err {err}"


@pytest.mark.parametrize("M,N,K,splits", [
    (160, 1280, 1280, 1), (160, 1280, 1280, 10), (160, 3840, 1280, 4), (160, 1280, 5120, 14), (160, 5120, 1280, 1),
    (5, 1280, 1280, 8), (1, 384, 384, 1), (37, 1000, 512, 3), (256, 768, 768, 2), (160, 51866, 1280, 1),
    (160, 1280, 1280, 5 | (64 << 8)), (160, 3840, 1280, 2 | (64 << 8)), (64, 1000, 512, 1 | (64 << 8)),
])
def test_skinny_gemm(lib, cuda_device, M, N, K, splits):
    """Decode-step GEMM (weight rows on the MMA M dimension) vs torch fp32; split-K slabs summed in order."""
    g = torch.Generator(device="cpu").manual_seed(M * 5 + N + K + splits)
    X = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda_device)
    ldo = (N + 7) // 8 * 8
    ref = X.float() @ W.float().T
    tol = 2e-3 * max(1.0, ref.abs().max().item())
    s = torch.cuda.current_stream().cuda_stream
    tile64 = splits & (64 << 8)  # bits 8.. of the `splits` argument select 64-row weight tiles (test hook)
    splits &= 0xff
    if splits > 1:
        out = torch.zeros((splits, M, ldo), dtype=torch.float32, device=cuda_device)
        _lib.check(lib.b200w_skinny_gemm(X.data_ptr(), W.data_ptr(), None, out.data_ptr(), M, N, K, K, K, ldo, 1, splits | tile64, s))
        torch.cuda.synchronize()
        got = out.sum(0)[:, :N]
        assert (got - ref).abs().max().item() <= tol
        return
    bias = torch.randn(N, generator=g).to(cuda_device)
    out = torch.zeros((M, ldo), dtype=torch.float32, device=cuda_device)
    _lib.check(lib.b200w_skinny_gemm(X.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, K, K, ldo, 1, 1 | tile64, s))
    torch.cuda.synchronize()
    assert (out[:, :N] - (ref + bias)).abs().max().item() <= tol
    out16 = torch.zeros((M, ldo), dtype=torch.bfloat16, device=cuda_device)
    _lib.check(lib.b200w_skinny_gemm(X.data_ptr(), W.data_ptr(), bias.data_ptr(), out16.data_ptr(), M, N, K, K, K, ldo, 4, 1 | tile64, s))
    torch.cuda.synchronize()
    refg = torch.nn.functional.gelu(ref + bias)
    assert (out16[:, :N].float() - refg).abs().max().item() <= 2e-2 * max(1.0, refg.abs().max().item())


@pytest.mark.parametrize("n_streams,H,L,layer,length,counts,pieces", [
    (2, 6, 2, 1, 1500, [5, 5], 3),              # two streams, beam 5: 12 items x 3 key-range pieces
    (32, 20, 1, 0, 1500, [5] * 32, 3),          # the bench shape: 640 items x 3 pieces over 296 CTAs
    (32, 20, 1, 0, 1500, [5] * 32, 1),          # the same, whole items (prompt-sized passes)
    (3, 4, 1, 0, 1500, [1, 8, 0], 4),           # greedy row, full group of 8, finished stream (skipped)
    (2, 8, 1, 0, 700, [3, 2], 2),               # ragged key count (6 tiles, last one partly out of range)
    (2, 8, 1, 0, 700, [3, 2], 6),               # one tile per piece; the last piece is mostly out of range
    (40, 8, 1, 0, 1500, [5] * 40, 2),           # 640 units > 296 CTAs with remainder
    (5, 6, 1, 0, 1500, [5] * 5, 5),             # 5 does not divide 12 tiles: falls back to whole items
])
def test_cross_attention_tc(lib, cuda_device, n_streams, H, L, layer, length, counts, pieces):
    """Decode-step cross attention (tcgen05) vs fp32 softmax(q k^T / 8) v on the same bf16 inputs."""
    g = torch.Generator(device="cpu").manual_seed(n_streams * 131 + H + length)
    d = H * 64
    kv = (torch.randn(n_streams, L, 2, H, length, 64, generator=g)).to(torch.bfloat16).to(cuda_device)
    rows_per = 8
    n_rows = n_streams * rows_per
    q = (torch.randn(n_rows, d, generator=g) * 1.5).to(torch.bfloat16).to(cuda_device)
    stream = torch.arange(n_streams, dtype=torch.int32).repeat_interleave(rows_per).to(cuda_device)
    live = [i for i, c in enumerate(counts) if c > 0]
    dead = [i for i, c in enumerate(counts) if c == 0]
    order = live + dead  # live groups first, as the engine orders them
    g_start = torch.tensor([i * rows_per for i in order], dtype=torch.int32, device=cuda_device)
    g_count = torch.tensor([counts[i] for i in order], dtype=torch.int32, device=cuda_device)
    n_live = torch.tensor([len(live)], dtype=torch.int32, device=cuda_device)
    out = torch.full((n_rows, d), 7.0, dtype=torch.bfloat16, device=cuda_device)
    scratch = torch.zeros(lib.b200w_cross_attention_tc_scratch_bytes(), dtype=torch.uint8, device=cuda_device)
    for _ in range(2):  # twice: the arrival counters of split items must reset themselves
        _lib.check(lib.b200w_cross_attention_tc(kv.data_ptr(), n_streams, L, layer, H, length, q.data_ptr(),
                                                stream.data_ptr(), g_start.data_ptr(), g_count.data_ptr(),
                                                n_live.data_ptr(), len(counts), out.data_ptr(), 0.125,
                                                scratch.data_ptr(), pieces, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    worst = 0.0
    for s_i, c in enumerate(counts):
        r0 = s_i * rows_per
        if c == 0:
            assert (out[r0:r0 + rows_per].float() == 7.0).all(), "rows of a finished stream must not be written"
            continue
        qs = q[r0:r0 + c].float().view(c, H, 64).transpose(0, 1)          # [H][c][64]
        k = kv[s_i, layer, 0].float()                                      # [H][len][64]
        v = kv[s_i, layer, 1].float()
        ref = torch.softmax(qs @ k.transpose(-1, -2) * 0.125, dim=-1) @ v  # [H][c][64]
        ref = ref.transpose(0, 1).reshape(c, d)
        worst = max(worst, (out[r0:r0 + c].float() - ref).abs().max().item())
        assert (out[r0 + c:r0 + rows_per].float() == 7.0).all(), "dead rows of a live group must not be written"
    # P is rounded to bf16 before P.V and the output is bf16: |v| <~ 4, so 2^-8 relative of a few units
    assert worst <= 3e-2, f"max abs err {worst}"
    # batch invariance, bit for bit: the first live stream ALONE in a launch gives exactly its rows of the batch
    s0 = live[0]
    out1 = torch.full((n_rows, d), 7.0, dtype=torch.bfloat16, device=cuda_device)
    one = torch.tensor([1], dtype=torch.int32, device=cuda_device)
    _lib.check(lib.b200w_cross_attention_tc(kv.data_ptr(), n_streams, L, layer, H, length, q.data_ptr(),
                                            stream.data_ptr(), g_start.data_ptr(), g_count.data_ptr(),
                                            one.data_ptr(), len(counts), out1.data_ptr(), 0.125,
                                            scratch.data_ptr(), pieces, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    r0, c = s0 * rows_per, counts[s0]
    assert torch.equal(out1[r0:r0 + c], out[r0:r0 + c]), "a stream's rows depend on what else is in the launch"


@pytest.mark.parametrize("n_streams,H,L,layer,length,counts", [
    (3, 6, 2, 1, 1500, [80, 5, 130]),              # > 128 rows in the last segment (two query tiles, ragged end)
    (32, 20, 1, 0, 1500, [70 + (i % 7) for i in range(32)]),   # the alignment pass of the bench workload
    (2, 4, 1, 0, 700, [9, 224]),                   # ragged key count; a long prompt
])
def test_cross_attention_flash(lib, cuda_device, n_streams, H, L, layer, length, counts):
    """Many-row cross attention (flash kernel over activation-row queries and cross K/V blocks) vs fp32."""
    g = torch.Generator(device="cpu").manual_seed(n_streams * 17 + H + length)
    d = H * 64
    kv = torch.randn(n_streams, L, 2, H, length, 64, generator=g).to(torch.bfloat16).to(cuda_device)
    n_rows = sum(counts)
    q = (torch.randn(n_rows, d, generator=g) * 1.5).to(torch.bfloat16).to(cuda_device)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    stream = torch.from_numpy(np.repeat(np.arange(n_streams, dtype=np.int32), counts)).to(cuda_device)
    seg_start = torch.from_numpy(starts).to(cuda_device)
    seg_count = torch.tensor(counts, dtype=torch.int32, device=cuda_device)
    out = torch.full((n_rows, d), 7.0, dtype=torch.bfloat16, device=cuda_device)
    _lib.check(lib.b200w_cross_attention_flash(kv.data_ptr(), n_streams, L, layer, H, length, q.data_ptr(), n_rows,
                                               stream.data_ptr(), seg_start.data_ptr(), seg_count.data_ptr(),
                                               n_streams, max(counts), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    worst = 0.0
    for s_i, c in enumerate(counts):
        r0 = int(starts[s_i])
        qs = q[r0:r0 + c].float().view(c, H, 64).transpose(0, 1)
        k = kv[s_i, layer, 0].float()
        v = kv[s_i, layer, 1].float()
        ref = (torch.softmax(qs @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).transpose(0, 1).reshape(c, d)
        worst = max(worst, (out[r0:r0 + c].float() - ref).abs().max().item())
    assert worst <= 3e-2, f"max abs err {worst}"
