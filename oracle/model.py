"""Oracle (test infrastructure): Whisper encoder / decoder on the CPU in PyTorch fp32.

Restates the CTranslate2 Whisper graph the reference reaches through faster-whisper
(third-party, un-vendored; call site /root/reference/asr_components.py:279-288; algorithm:
SURVEY.md Appendix A.5 == OpenAI whisper/model.py).  PARITY UNPINNED by the reference (it ships
no vectors); cross-checked against transformers' WhisperModel with shared random weights in
tests/test_oracle_model.py.

Two precisions:
  "fp32"  everything in float32 (weights are whatever the state dict holds).
  "bf16"  float32 arithmetic, but every tensor that a bf16 inference engine materialises between
          ops (LayerNorm / attention / GELU outputs, Q K V, softmax probabilities of the encoder,
          encoder output, cross K/V) is rounded to bfloat16 at that point.  This is what
          faster-whisper's compute_type="bfloat16" stores between CTranslate2 ops, and it is the
          precision contract the token-exactness tests hold the CUDA path to.  The residual stream
          and LayerNorm statistics stay in fp32.
  "int8"  CPU-baseline variant only (bench.py), never a parity reference: every Linear layer (and the tied output
          projection) runs as CTranslate2's int8 CPU path does -- weights quantised to int8 per output
          row, activations quantised dynamically per call, s8 x s8 -> s32 GEMM (torch's quantized.linear_dynamic on
          the x86 / fbgemm engine) -- with conv, LayerNorm, softmax and GELU in fp32.
State-dict keys follow OpenAI Whisper (encoder.blocks.0.attn.query.weight, ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class Dims:
    n_mels: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int
    n_audio_ctx: int = 1500
    n_text_ctx: int = 448


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """OpenAI whisper sinusoids(): encoder positional embedding."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


def _bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class OracleWhisper:
    def __init__(self, sd: Dict[str, torch.Tensor], dims: Dims, precision: str = "fp32"):
        assert precision in ("fp32", "bf16", "int8")
        self.sd = {k: v.to(torch.float32) for k, v in sd.items()}
        self.dims = dims
        self.precision = precision
        self.r = _bf if precision == "bf16" else (lambda x: x)
        self._packed = {}  # int8: prefix -> prepacked (per-row int8 weight, fp32 bias)

    def _int8_linear(self, x, prefix, bias=True):
        pk = self._packed.get((prefix, bias))
        if pk is None:
            w = self.sd[prefix + ".weight"]
            b = self.sd.get(prefix + ".bias") if bias else None
            scales = (w.abs().amax(dim=1) / 127.0).clamp_min(1e-12)
            qw = torch.quantize_per_channel(w, scales, torch.zeros(w.shape[0], dtype=torch.int64), 0, torch.qint8)
            pk = torch.ops.quantized.linear_prepack(qw, b)
            self._packed[(prefix, bias)] = pk
        shp = x.shape
        y = torch.ops.quantized.linear_dynamic(x.reshape(-1, shp[-1]).contiguous(), pk)
        return y.reshape(*shp[:-1], y.shape[-1])

    # ------------------------------------------------------------------ helpers
    def _lin(self, x, prefix, bias=True):
        if self.precision == "int8":
            return self._int8_linear(x, prefix, bias)
        w = self.sd[prefix + ".weight"]
        b = self.sd.get(prefix + ".bias") if bias else None
        return F.linear(x, w, b)

    def _ln(self, x, prefix):
        d = x.shape[-1]
        return F.layer_norm(x, (d,), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], 1e-5)

    def _enc_attention(self, q, k, v, n_head):
        # q,k,v: [B,T,d]
        B, T, d = q.shape
        q = q.view(B, T, n_head, 64).transpose(1, 2)
        k = k.view(B, T, n_head, 64).transpose(1, 2)
        v = v.view(B, T, n_head, 64).transpose(1, 2)
        if self.precision == "fp32":
            w = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1)
            o = w @ v
        else:
            # same schedule as the CUDA kernel: 64-key blocks, running max, probabilities rounded
            # to bf16 (relative to the running max of that moment) before the P.V product, row
            # sums taken before the rounding
            c = 0.125 * 1.4426950408889634
            m = torch.full((B, n_head, T, 1), -float("inf"))
            l = torch.zeros((B, n_head, T, 1))
            o = torch.zeros((B, n_head, T, 64))
            for j0 in range(0, T, 64):
                s = q @ k[:, :, j0:j0 + 64].transpose(-1, -2)
                m_new = torch.maximum(m, s.max(dim=-1, keepdim=True).values * c)
                alpha = torch.exp2(m - m_new)
                p = torch.exp2(s * c - m_new)
                l = l * alpha + p.sum(dim=-1, keepdim=True)
                o = o * alpha + _bf(p) @ v[:, :, j0:j0 + 64]
                m = m_new
            o = o / l
        return o.transpose(1, 2).reshape(B, T, d)

    # ------------------------------------------------------------------ encoder
    def encode(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, n_mels, 3000] float32 -> [B, 1500, d]."""
        r, D = self.r, self.dims
        x = r(mel.to(torch.float32))
        x = r(F.gelu(F.conv1d(x, self.sd["encoder.conv1.weight"], self.sd["encoder.conv1.bias"], padding=1)))
        x = F.gelu(F.conv1d(x, self.sd["encoder.conv2.weight"], self.sd["encoder.conv2.bias"], stride=2, padding=1))
        x = x.permute(0, 2, 1) + self.sd["encoder.positional_embedding"]
        for i in range(D.n_audio_layer):
            p = f"encoder.blocks.{i}"
            h = r(self._ln(x, p + ".attn_ln"))
            q = r(self._lin(h, p + ".attn.query"))
            k = r(self._lin(h, p + ".attn.key", bias=False))
            v = r(self._lin(h, p + ".attn.value"))
            a = r(self._enc_attention(q, k, v, D.n_audio_head))
            x = x + self._lin(a, p + ".attn.out")
            h = r(self._ln(x, p + ".mlp_ln"))
            h = r(F.gelu(self._lin(h, p + ".mlp.0")))
            x = x + self._lin(h, p + ".mlp.2")
        return r(self._ln(x, "encoder.ln_post"))

    # ------------------------------------------------------------------ decoder
    def cross_kv(self, enc: torch.Tensor):
        """Per decoder layer (K, V) of the encoder output, computed once per chunk."""
        out = []
        for i in range(self.dims.n_text_layer):
            p = f"decoder.blocks.{i}.cross_attn"
            out.append((self.r(self._lin(enc, p + ".key", bias=False)), self.r(self._lin(enc, p + ".value"))))
        return out

    def new_state(self, n_rows: int):
        return {"k": [None] * self.dims.n_text_layer, "v": [None] * self.dims.n_text_layer, "len": 0,
                "rows": n_rows}

    @staticmethod
    def reorder_state(state, parents: List[int]):
        idx = torch.tensor(parents, dtype=torch.long)
        state["k"] = [t.index_select(0, idx) if t is not None else None for t in state["k"]]
        state["v"] = [t.index_select(0, idx) if t is not None else None for t in state["v"]]
        state["rows"] = len(parents)
        return state

    def _attend(self, q, k, v, n_head, causal_offset: Optional[int], want_weights=False):
        # q [R,Tq,d]; k,v [R,Tk,d]
        R, Tq, d = q.shape
        Tk = k.shape[1]
        q = q.view(R, Tq, n_head, 64).transpose(1, 2) * 0.125
        k = k.view(R, Tk, n_head, 64).transpose(1, 2)
        v = v.view(R, Tk, n_head, 64).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        if causal_offset is not None:
            qi = torch.arange(Tq)[:, None] + causal_offset
            kj = torch.arange(Tk)[None, :]
            s = s.masked_fill(kj > qi, -float("inf"))
        w = torch.softmax(s, dim=-1)
        o = (w @ v).transpose(1, 2).reshape(R, Tq, d)
        return (o, w) if want_weights else (o, None)

    def decode_step(self, tokens: torch.Tensor, state, cross, row_stream: List[int], want_cross_weights=False):
        """tokens [R, Tq] appended at positions state['len']...; returns hidden [R,Tq,d] (pre-ln)
        and optionally the cross-attention probabilities per layer [R, H, Tq, 1500]."""
        r, D = self.r, self.dims
        off = state["len"]
        Tq = tokens.shape[1]
        x = self.sd["decoder.token_embedding.weight"][tokens] + self.sd["decoder.positional_embedding"][off:off + Tq]
        sidx = torch.tensor(row_stream, dtype=torch.long)
        cross_w = []
        for i in range(D.n_text_layer):
            p = f"decoder.blocks.{i}"
            h = r(self._ln(x, p + ".attn_ln"))
            q = r(self._lin(h, p + ".attn.query"))
            k = r(self._lin(h, p + ".attn.key", bias=False))
            v = r(self._lin(h, p + ".attn.value"))
            state["k"][i] = k if state["k"][i] is None else torch.cat([state["k"][i], k], dim=1)
            state["v"][i] = v if state["v"][i] is None else torch.cat([state["v"][i], v], dim=1)
            a, _ = self._attend(q, state["k"][i], state["v"][i], D.n_text_head, off)
            x = x + self._lin(r(a), p + ".attn.out")
            h = r(self._ln(x, p + ".cross_attn_ln"))
            q = r(self._lin(h, p + ".cross_attn.query"))
            ck, cv = cross[i]
            a, w = self._attend(q, ck.index_select(0, sidx), cv.index_select(0, sidx), D.n_text_head, None,
                                want_weights=want_cross_weights)
            if want_cross_weights:
                cross_w.append(w)
            x = x + self._lin(r(a), p + ".cross_attn.out")
            h = r(self._ln(x, p + ".mlp_ln"))
            h = r(F.gelu(self._lin(h, p + ".mlp.0")))
            x = x + self._lin(h, p + ".mlp.2")
        state["len"] = off + Tq
        return x, cross_w

    def logits(self, hidden: torch.Tensor) -> torch.Tensor:
        h = self.r(self._ln(hidden, "decoder.ln"))
        if self.precision == "int8":
            return self._int8_linear(h, "decoder.token_embedding", bias=False)
        return h @ self.sd["decoder.token_embedding.weight"].T

    def teacher_forced_logits(self, tokens: List[int], enc_row: torch.Tensor, want_cross_weights=False):
        """Full-sequence pass for one sequence; returns logits [T, V] (+ cross weights)."""
        cross = self.cross_kv(enc_row[None])
        st = self.new_state(1)
        hid, cw = self.decode_step(torch.tensor([tokens], dtype=torch.long), st, cross, [0], want_cross_weights)
        return self.logits(hid)[0], cw


def convert_hf_state_dict(hf_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """transformers WhisperModel state dict -> OpenAI Whisper keys (for the cross-check test)."""
    out = {}
    ren = [("model.", ""), ("layers.", "blocks."), ("self_attn_layer_norm", "attn_ln"),
           ("encoder_attn_layer_norm", "cross_attn_ln"), ("final_layer_norm", "mlp_ln"),
           ("self_attn.q_proj", "attn.query"), ("self_attn.k_proj", "attn.key"),
           ("self_attn.v_proj", "attn.value"), ("self_attn.out_proj", "attn.out"),
           ("encoder_attn.q_proj", "cross_attn.query"), ("encoder_attn.k_proj", "cross_attn.key"),
           ("encoder_attn.v_proj", "cross_attn.value"), ("encoder_attn.out_proj", "cross_attn.out"),
           ("fc1", "mlp.0"), ("fc2", "mlp.2"), ("embed_positions.weight", "positional_embedding"),
           ("embed_tokens", "token_embedding"), ("encoder.layer_norm", "encoder.ln_post"),
           ("decoder.layer_norm", "decoder.ln")]
    for k, v in hf_sd.items():
        if k.startswith("proj_out"):
            continue
        nk = k
        for a, b in ren:
            nk = nk.replace(a, b)
        out[nk] = v.detach().to(torch.float32)
    return out
