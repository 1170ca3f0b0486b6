"""Model shapes, seeded synthetic weights, and packing of a Whisper state dict into the device
layout libb200whisper.so expects.

Canonical host format: OpenAI-Whisper state-dict keys (``encoder.blocks.0.attn.query.weight`` ...),
fp32 tensors whose matrix / embedding / conv values are already bf16-representable, so the CUDA
engine (bf16 operands) and the CPU oracle (fp32 arithmetic) see identical numbers.

Engine tensor names (device):
  enc.conv1.w bf16 [d][3*128] tap-major, channels zero-padded to 128     enc.conv1.b f32 [d]
  enc.conv2.w bf16 [d][3*d]  tap-major                                   enc.conv2.b f32 [d]
  enc.pos f32 [1500][d]
  enc.L{i}.{ln1,ln2}.{g,b} f32 [d]    enc.L{i}.qkv.w bf16 [3d][d] (+.b, key bias = 0)
  enc.L{i}.{out,fc1,fc2}.{w,b}        enc.ln_post.{g,b}
  dec.tok_emb bf16 [V][d] (also the logits matrix)    dec.pos f32 [448][d]
  dec.L{i}.{ln1,ln2,ln3}  dec.L{i}.{qkv,out,cq,cout,fc1,fc2}.{w,b}
  dec.cross_kv.{w,b}  bf16 [L*2d][d]: per layer [key ; value] of the cross attention, all layers
                      stacked so that the per-chunk cross K/V is ONE GEMM
  dec.ln.{g,b}
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch


@dataclass(frozen=True)
class ModelDims:
    n_mels: int
    d_model: int
    n_heads: int
    n_enc_layers: int
    n_dec_layers: int
    n_vocab: int
    n_audio_ctx: int = 1500
    n_text_ctx: int = 448

    @property
    def multilingual(self) -> bool:
        return self.n_vocab >= 51865


MODEL_DIMS: Dict[str, ModelDims] = {
    # "micro": test-only shape (not an OpenAI size) that keeps CPU-oracle tests fast
    "micro": ModelDims(80, 128, 2, 2, 2, 51865),
    "tiny": ModelDims(80, 384, 6, 4, 4, 51865),
    "base": ModelDims(80, 512, 8, 6, 6, 51865),
    "small": ModelDims(80, 768, 12, 12, 12, 51865),
    "medium": ModelDims(80, 1024, 16, 24, 24, 51865),
    "large-v2": ModelDims(80, 1280, 20, 32, 32, 51865),
    "large-v3": ModelDims(128, 1280, 20, 32, 32, 51866),
}
MODEL_DIMS["large"] = MODEL_DIMS["large-v3"]

# (layer, head) pairs whose cross-attention tracks word timing (OpenAI _ALIGNMENT_HEADS, what the
# CTranslate2 conversions store as `alignment_heads` in config.json).
ALIGNMENT_HEADS: Dict[str, List[Tuple[int, int]]] = {
    "tiny": [(2, 2), (3, 0), (3, 2), (3, 3), (3, 4), (3, 5)],
    "base": [(3, 1), (4, 2), (4, 3), (4, 7), (5, 1), (5, 2), (5, 4), (5, 6)],
    "small": [(5, 3), (5, 9), (8, 0), (8, 4), (8, 7), (8, 8), (9, 0), (9, 7), (9, 9), (10, 5)],
    "medium": [(13, 15), (15, 4), (15, 15), (16, 1), (20, 0), (23, 4)],
    "large-v2": [(10, 12), (13, 17), (16, 11), (16, 12), (16, 13), (17, 15), (17, 16), (18, 4),
                 (18, 11), (18, 19), (19, 11), (21, 2), (21, 3), (22, 3), (22, 9), (22, 12),
                 (23, 5), (23, 7), (23, 13), (25, 5), (26, 1), (26, 12), (27, 15)],
    "large-v3": [(7, 0), (10, 17), (12, 18), (13, 12), (16, 1), (17, 14), (19, 11), (21, 4),
                 (24, 1), (25, 6)],
    "micro": [(1, 0), (1, 1)],
}
ALIGNMENT_HEADS["large"] = ALIGNMENT_HEADS["large-v3"]


def default_alignment_heads(dims: ModelDims) -> List[Tuple[int, int]]:
    """CTranslate2's default when config.json has none: every head of the upper half of layers."""
    return [(l, h) for l in range(dims.n_dec_layers // 2, dims.n_dec_layers) for h in range(dims.n_heads)]


def alignment_heads_for(dims: ModelDims) -> List[Tuple[int, int]]:
    """The (layer, head) list a model of these dimensions aligns with: OpenAI's list for the known sizes,
    CTranslate2's default otherwise."""
    for name, md in MODEL_DIMS.items():
        if md == dims and name in ALIGNMENT_HEADS:
            return list(ALIGNMENT_HEADS[name])
    return default_alignment_heads(dims)


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def synthetic_state_dict(dims: ModelDims, seed: int = 0, device: str = "cpu",
                         peaked: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded random weights of the exact architecture (there are no checkpoints offline).

    Variance-scaled so activations stay O(1) through the depth.  Matrices / embeddings are rounded
    to bf16-representable values.

    peaked=False: plain random model.  Its next-token distribution is nearly flat (top-2 logit gaps
      of a fraction of a nat), which no bf16 engine can reproduce token-for-token; it is the right
      model for tolerance tests on activations and logits.
    peaked=True (default): the same random model plus a "transcript circuit" (see
      _install_transcript_circuit) that makes the decoder confident like a trained model: it walks
      through one of 51 synthetic transcripts (timestamp, words, timestamp pairs, ..., EOT), chosen
      at the first sampled position from the audio-dependent context.  This gives realistic
      segment structure and decode lengths for end-to-end tests and benchmarks.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    d, L_e, L_d = dims.d_model, dims.n_enc_layers, dims.n_dec_layers
    bgain = 0.5 if peaked else 2.0
    cgain = (1.0 if d >= 384 else 0.5) if peaked else bgain  # attention branches of the decoder
    sd: Dict[str, torch.Tensor] = {}

    def rnd(*shape, std):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std

    def lin(prefix, n_out, n_in, gain=1.0, bias=True):
        sd[prefix + ".weight"] = _bf16_round(rnd(n_out, n_in, std=gain / math.sqrt(n_in)))
        if bias:
            sd[prefix + ".bias"] = rnd(n_out, std=0.02)

    def ln(prefix):
        sd[prefix + ".weight"] = 1.0 + rnd(d, std=0.05)
        sd[prefix + ".bias"] = rnd(d, std=0.02)

    sd["encoder.conv1.weight"] = _bf16_round(rnd(d, dims.n_mels, 3, std=1.0 / math.sqrt(3 * dims.n_mels)))
    sd["encoder.conv1.bias"] = rnd(d, std=0.02)
    sd["encoder.conv2.weight"] = _bf16_round(rnd(d, d, 3, std=1.0 / math.sqrt(3 * d)))
    sd["encoder.conv2.bias"] = rnd(d, std=0.02)
    sd["encoder.positional_embedding"] = sinusoids(dims.n_audio_ctx, d).to(device)
    for i in range(L_e):
        p = f"encoder.blocks.{i}"
        ln(p + ".attn_ln")
        lin(p + ".attn.query", d, d, gain=2.0)
        lin(p + ".attn.key", d, d, bias=False)
        lin(p + ".attn.value", d, d)
        lin(p + ".attn.out", d, d, gain=bgain / math.sqrt(L_e))
        ln(p + ".mlp_ln")
        lin(p + ".mlp.0", 4 * d, d)
        lin(p + ".mlp.2", d, 4 * d, gain=bgain / math.sqrt(L_e))
    ln("encoder.ln_post")
    sd["decoder.token_embedding.weight"] = _bf16_round(rnd(dims.n_vocab, d, std=1.5 / math.sqrt(d)))
    sd["decoder.positional_embedding"] = rnd(dims.n_text_ctx, d, std=0.1)
    for i in range(L_d):
        p = f"decoder.blocks.{i}"
        ln(p + ".attn_ln")
        lin(p + ".attn.query", d, d, gain=2.0)
        lin(p + ".attn.key", d, d, bias=False)
        lin(p + ".attn.value", d, d)
        lin(p + ".attn.out", d, d, gain=cgain / math.sqrt(L_d))
        ln(p + ".cross_attn_ln")
        lin(p + ".cross_attn.query", d, d, gain=3.0)
        lin(p + ".cross_attn.key", d, d, bias=False)
        lin(p + ".cross_attn.value", d, d)
        lin(p + ".cross_attn.out", d, d, gain=cgain / math.sqrt(L_d))
        ln(p + ".mlp_ln")
        lin(p + ".mlp.0", 4 * d, d)
        lin(p + ".mlp.2", d, 4 * d, gain=bgain / math.sqrt(L_d))
    ln("decoder.ln")
    if peaked:
        _install_transcript_circuit(sd, dims, g, device)
        _install_alignment_heads(sd, dims, g, device)
    return sd


def special_token_ids(n_vocab: int) -> Dict[str, int]:
    """Special-token layout of the multilingual vocabularies."""
    s = 1 if n_vocab >= 51866 else 0
    return dict(eot=50257, sot=50258, lang0=50259, n_lang=99 + s, translate=50358 + s, transcribe=50359 + s,
                sot_lm=50360 + s, sot_prev=50361 + s, no_speech=50362 + s, no_timestamps=50363 + s,
                timestamp_begin=50364 + s)


def synthetic_transcripts(n_vocab: int, seed: int = 0) -> List[List[int]]:
    """The 51 token sequences (without EOT) the peaked synthetic model can emit: transcript j
    starts at timestamp j (0.00 .. 1.00 s) and alternates ~1 s segments of 3-8 word pieces with
    ~1 s pauses; its timestamps are tb + j + 51 m, so no token occurs twice anywhere."""
    sp = special_token_ids(n_vocab)
    tb = sp["timestamp_begin"]
    rng = torch.Generator().manual_seed(1000 + seed)
    words = (torch.randperm(50000 - 256, generator=rng) + 256).tolist()
    lens = torch.randint(3, 9, (51 * 16,), generator=rng).tolist()
    out, wi, li = [], 0, 0
    for j in range(51):
        seq, m = [], 0
        while tb + j + 51 * (m + 1) < n_vocab and m < 28:
            seq.append(tb + j + 51 * m)          # segment start
            k = lens[li]
            li += 1
            seq.extend(words[wi:wi + k])
            wi += k
            seq.append(tb + j + 51 * (m + 1))    # segment end
            m += 2
        out.append(seq)
    return out


def _install_transcript_circuit(sd: Dict[str, torch.Tensor], dims: ModelDims, g, device) -> None:
    """Make the random decoder confident without training.

    Token embeddings are laid out along orbits of a random rotation R: E[next] = R E[cur] for
    consecutive tokens of each synthetic transcript.  The last decoder MLP is set to the exact
    linear map u -> gamma R u (GELU(z) - GELU(-z) = z), so the final hidden state points at the
    embedding of the successor token with a margin far above bf16 noise.  Every transcript's last
    timestamp shares a component with R^-1 E[eot] so that EOT follows it.  All other layers stay
    random and still move the logits (they decide the first timestamp, i.e. which transcript)."""
    d, V, L = dims.d_model, dims.n_vocab, dims.n_dec_layers
    sp = special_token_ids(V)
    s = 0.5 * math.sqrt(d)
    gamma, alpha = 3.0, 0.7
    beta = math.sqrt(1.0 - alpha * alpha)
    q, r = torch.linalg.qr(torch.randn(d, d, generator=g, device=device, dtype=torch.float64))
    R = (q * torch.sign(torch.diagonal(r))).to(torch.float32)

    def unit(n):
        v = torch.randn(n, d, generator=g, device=device, dtype=torch.float32)
        return v / v.norm(dim=1, keepdim=True)

    E = unit(V) * (0.5 * s)                       # background tokens: half the norm
    E[sp["eot"]] = unit(1)[0] * s
    for name in ("sot", "translate", "transcribe", "sot_prev", "no_timestamps"):
        E[sp[name]] = unit(1)[0] * s
    e_end = R.T @ E[sp["eot"]]                    # R e_end = E[eot]
    for seq in synthetic_transcripts(V, 0):
        n_j = unit(1)[0]
        n_j = n_j - (n_j @ e_end) / (e_end @ e_end) * e_end
        n_j = n_j / n_j.norm()
        e = alpha * e_end + beta * s * n_j        # last timestamp of the transcript
        for tok in reversed(seq):                 # walk the orbit backwards: E[prev] = R^T E[cur]
            E[tok] = e
            e = R.T @ e
    sd["decoder.token_embedding.weight"] = _bf16_round(E)
    p = f"decoder.blocks.{L - 1}"
    fc1 = torch.zeros(4 * d, d, device=device)
    fc1[:d] = torch.eye(d, device=device)
    fc1[d:2 * d] = -torch.eye(d, device=device)
    fc2 = torch.zeros(d, 4 * d, device=device)
    fc2[:, :d] = gamma * R
    fc2[:, d:2 * d] = -gamma * R
    sd[p + ".mlp.0.weight"] = _bf16_round(fc1)
    sd[p + ".mlp.0.bias"] = torch.zeros(4 * d, device=device)
    sd[p + ".mlp.2.weight"] = _bf16_round(fc2)
    sd[p + ".mlp.2.bias"] = torch.zeros(d, device=device)
    for n in (p + ".mlp_ln", "decoder.ln"):
        sd[n + ".weight"] = torch.ones(d, device=device)
        sd[n + ".bias"] = torch.zeros(d, device=device)


ALIGN_STRIDE = 12       # encoder frames (20 ms each) the synthetic alignment heads advance per decoder position
ALIGN_CODE_AMPLITUDE = 2.0


def _install_alignment_heads(sd: Dict[str, torch.Tensor], dims: ModelDims, g, device) -> None:
    """Give the alignment heads the sharply peaked, monotone cross-attention a trained Whisper has.

    A trained model's alignment heads attend, at decoder position i, to the audio frames where token i is
    spoken; a random head attends almost uniformly, its DTW cost surface is flat and the word-timing path
    is decided by rounding noise.  Here the heads read POSITION on both sides: the encoder output still
    carries its sinusoidal position code (residual stream), the decoder's positional embedding gets a
    matching code of frame f(i) = ALIGN_STRIDE * i in a 64-dimensional subspace, the head's key projection
    selects 32 (sin, cos) pairs of the encoder code and its query projection reads the decoder code:
        q_i . k_t  ~  sum_j cos(w_j (t - f(i)))      -> a peak at t = f(i) a few frames wide, side lobes ~ sqrt(32).
    Values and output projections stay random, so through these heads every decoder position reads the LOCAL
    audio content around frame f(i) instead of a time average: the logits (and the first sampled timestamp,
    i.e. which synthetic transcript is emitted) depend on the audio."""
    d, half = dims.d_model, dims.d_model // 2
    heads = alignment_heads_for(dims)
    # 32 encoder frequencies w_m = 10000^(-m/(half-1)) between ~0.3 and ~3e-3 rad/frame: the peak is ~+-4 frames
    # wide, so it survives the width-7 median filter of the word-timing pass
    m_lo, m_hi = int(0.12 * (half - 1)), int(0.62 * (half - 1))
    m_idx = torch.linspace(m_lo, m_hi, 32).round().long().unique()
    assert len(m_idx) == 32, "not enough distinct position frequencies"
    inc = math.log(10000.0) / (half - 1)
    w = torch.exp(-inc * m_idx.to(torch.float32))                          # [32]
    pos = ALIGN_STRIDE * torch.arange(dims.n_text_ctx, dtype=torch.float32)
    code = torch.stack([torch.sin(pos[:, None] * w), torch.cos(pos[:, None] * w)], dim=2).reshape(dims.n_text_ctx, 64)
    q_, r_ = torch.linalg.qr(torch.randn(d, 64, generator=g, device=device, dtype=torch.float64))
    P = (q_ * torch.sign(torch.diagonal(r_))).to(torch.float32)           # d x 64, orthonormal columns
    sd["decoder.positional_embedding"] = sd["decoder.positional_embedding"] + ALIGN_CODE_AMPLITUDE * (code.to(device) @ P.T)
    for (l, h) in heads:
        p = f"decoder.blocks.{l}.cross_attn"
        wk = sd[p + ".key.weight"]
        wq = sd[p + ".query.weight"]
        rows = slice(h * 64, (h + 1) * 64)
        wk[rows] = 0.0
        for j, m in enumerate(m_idx.tolist()):
            wk[h * 64 + 2 * j, m] = 1.0            # sin(w_m t) of the encoder position code
            wk[h * 64 + 2 * j + 1, half + m] = 1.0  # cos(w_m t)
        wq[rows] = _bf16_round(2.0 * P.T)
        sd[p + ".query.bias"][rows] = 0.0


def dims_from_state_dict(sd: Dict[str, torch.Tensor]) -> ModelDims:
    d = sd["encoder.conv1.weight"].shape[0]
    n_enc = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.blocks."))
    n_dec = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.blocks."))
    return ModelDims(sd["encoder.conv1.weight"].shape[1], d, d // 64, n_enc, n_dec,
                     sd["decoder.token_embedding.weight"].shape[0],
                     sd["encoder.positional_embedding"].shape[0],
                     sd["decoder.positional_embedding"].shape[0])


def pack_for_engine(sd: Dict[str, torch.Tensor], dims: ModelDims, device) -> Dict[str, torch.Tensor]:
    """OpenAI-keyed state dict -> {engine name: device tensor} (bf16 matrices, fp32 vectors)."""
    d = dims.d_model
    out: Dict[str, torch.Tensor] = {}

    def f32(t):
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    def b16(t):
        return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()

    def lin(dst, src, has_bias=True):
        out[dst + ".w"] = b16(sd[src + ".weight"])
        out[dst + ".b"] = f32(sd[src + ".bias"]) if has_bias else torch.zeros(
            sd[src + ".weight"].shape[0], dtype=torch.float32, device=device)

    def norm(dst, src):
        out[dst + ".g"] = f32(sd[src + ".weight"])
        out[dst + ".b"] = f32(sd[src + ".bias"])

    def fused_qkv(dst, src):
        q, k, v = (sd[f"{src}.{n}.weight"] for n in ("query", "key", "value"))
        out[dst + ".w"] = b16(torch.cat([q, k, v], dim=0))
        out[dst + ".b"] = f32(torch.cat([sd[src + ".query.bias"], torch.zeros_like(sd[src + ".query.bias"]),
                                         sd[src + ".value.bias"]]))

    w1 = sd["encoder.conv1.weight"].permute(0, 2, 1)  # [d, 3, n_mels]
    w1p = torch.zeros((d, 3, 128), dtype=w1.dtype, device=w1.device)
    w1p[:, :, : dims.n_mels] = w1
    out["enc.conv1.w"] = b16(w1p.reshape(d, 3 * 128))
    out["enc.conv1.b"] = f32(sd["encoder.conv1.bias"])
    out["enc.conv2.w"] = b16(sd["encoder.conv2.weight"].permute(0, 2, 1).reshape(d, 3 * d))
    out["enc.conv2.b"] = f32(sd["encoder.conv2.bias"])
    out["enc.pos"] = f32(sd["encoder.positional_embedding"])
    for i in range(dims.n_enc_layers):
        s, t = f"encoder.blocks.{i}", f"enc.L{i}"
        norm(t + ".ln1", s + ".attn_ln")
        fused_qkv(t + ".qkv", s + ".attn")
        lin(t + ".out", s + ".attn.out")
        norm(t + ".ln2", s + ".mlp_ln")
        lin(t + ".fc1", s + ".mlp.0")
        lin(t + ".fc2", s + ".mlp.2")
    norm("enc.ln_post", "encoder.ln_post")
    out["dec.tok_emb"] = b16(sd["decoder.token_embedding.weight"])
    out["dec.pos"] = f32(sd["decoder.positional_embedding"])
    ckw, ckb = [], []
    for i in range(dims.n_dec_layers):
        s, t = f"decoder.blocks.{i}", f"dec.L{i}"
        norm(t + ".ln1", s + ".attn_ln")
        fused_qkv(t + ".qkv", s + ".attn")
        lin(t + ".out", s + ".attn.out")
        norm(t + ".ln2", s + ".cross_attn_ln")
        lin(t + ".cq", s + ".cross_attn.query")
        lin(t + ".cout", s + ".cross_attn.out")
        norm(t + ".ln3", s + ".mlp_ln")
        lin(t + ".fc1", s + ".mlp.0")
        lin(t + ".fc2", s + ".mlp.2")
        ckw += [sd[s + ".cross_attn.key.weight"], sd[s + ".cross_attn.value.weight"]]
        ckb += [torch.zeros_like(sd[s + ".cross_attn.value.bias"]), sd[s + ".cross_attn.value.bias"]]
    out["dec.cross_kv.w"] = b16(torch.cat(ckw, dim=0))
    out["dec.cross_kv.b"] = f32(torch.cat(ckb, dim=0))
    norm("dec.ln", "decoder.ln")
    return out
