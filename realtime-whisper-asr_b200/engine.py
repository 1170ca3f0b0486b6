"""Python handle on the C++ engine in libb200whisper.so (encoder, cross K/V, decode, word timing).

PyTorch only provides device buffers and the CUDA stream; all compute is in the shared library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .weights import ModelDims


@dataclass
class DecodeOptions:
    beam_size: int = 5
    patience: float = 1.0
    length_penalty: float = 1.0
    max_new_tokens: int = 0
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp_index: int = 50
    eot: int = 50257
    sot: int = 50258
    no_speech: int = 50362
    no_timestamps: int = 50363
    timestamp_begin: int = 50364
    suppress_ids: Sequence[int] = field(default_factory=list)
    blank_ids: Sequence[int] = field(default_factory=list)
    temperature: float = 0.0   # > 0: best_of independent samples instead of greedy / beam search
    best_of: int = 1
    seed: int = 0


@dataclass
class DecodeResult:
    tokens: List[int]
    sum_logprob: float
    no_speech_prob: float


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p_i32(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _p_f32(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class WhisperEngine:
    def __init__(self, dims: ModelDims, packed: Dict[str, torch.Tensor], max_batch: int = 8,
                 max_beam: int = 5, device=None):
        if not torch.cuda.is_available():
            raise _lib.B200WError("CUDA device required: the B200 backend has no CPU fallback")
        self.lib = _lib.load()
        self.dims = dims
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.max_batch = max_batch
        self.max_beam = max_beam
        self._weights = packed  # keep the device tensors alive
        cfg = _lib.ModelConfig(dims.n_mels, dims.d_model, dims.n_heads, dims.n_enc_layers, dims.n_dec_layers,
                               dims.n_vocab, dims.n_audio_ctx, dims.n_text_ctx, max_batch, max_beam)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200w_engine_create(C.byref(cfg), C.byref(h)))
            self._h = h
            for name, t in packed.items():
                assert t.is_cuda and t.is_contiguous(), name
                dt = {torch.float32: 0, torch.bfloat16: 1}[t.dtype]
                _lib.check(self.lib.b200w_engine_set_weight(h, name.encode(), t.data_ptr(), t.numel(), dt))
            _lib.check(self.lib.b200w_engine_finalize(h))
        self.ldv = (dims.n_vocab + 7) // 8 * 8

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200w_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ encoder
    def encode(self, windows: torch.Tensor) -> torch.Tensor:
        """windows bf16 [B, 3000, 128] -> encoder output bf16 [B, 1500, d]."""
        assert windows.dtype == torch.bfloat16 and windows.shape[1:] == (3000, 128) and windows.is_contiguous()
        B = windows.shape[0]
        out = torch.empty((B, self.dims.n_audio_ctx, self.dims.d_model), dtype=torch.bfloat16, device=self.device)
        for b0 in range(0, B, self.max_batch):
            nb = min(self.max_batch, B - b0)
            _lib.check(self.lib.b200w_encode(self._h, windows[b0:b0 + nb].data_ptr(), nb,
                                             out[b0:b0 + nb].data_ptr(), self._stream()))
        return out

    def cross_kv(self, enc_out: torch.Tensor) -> torch.Tensor:
        """bf16 [B, 1500, d] -> bf16 [B, L_dec, 2, heads, 1500, 64] (head-major: one head's keys contiguous)."""
        B = enc_out.shape[0]
        assert B <= self.max_batch
        out = torch.empty((B, self.dims.n_dec_layers, 2, self.dims.n_heads, self.dims.n_audio_ctx, 64),
                          dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.b200w_cross_kv(self._h, enc_out.data_ptr(), B, out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------ decoder
    def _opts(self, o: DecodeOptions):
        sup = _i32(list(o.suppress_ids))
        blk = _i32(list(o.blank_ids))
        c = _lib.DecodeOpts(o.beam_size, o.patience, o.length_penalty, o.max_new_tokens,
                            int(o.suppress_blank), int(o.without_timestamps), o.max_initial_timestamp_index,
                            o.eot, o.sot, o.no_speech, o.no_timestamps, o.timestamp_begin,
                            _p_i32(sup), len(sup), _p_i32(blk), len(blk),
                            float(o.temperature), int(o.best_of), int(o.seed) & 0xFFFFFFFF)
        return c, (sup, blk)

    def decode(self, cross_kv: torch.Tensor, prompts: Sequence[Sequence[int]], opts: DecodeOptions) -> List[DecodeResult]:
        B = len(prompts)
        assert cross_kv.shape[0] == B and B <= self.max_batch
        flat = _i32([t for p in prompts for t in p])
        lens = _i32([len(p) for p in prompts])
        max_out = self.dims.n_text_ctx
        toks = np.zeros((B, max_out), dtype=np.int32)
        ntok = np.zeros(B, dtype=np.int32)
        slp = np.zeros(B, dtype=np.float32)
        nsp = np.zeros(B, dtype=np.float32)
        c, keep = self._opts(opts)
        _lib.check(self.lib.b200w_decode(self._h, cross_kv.data_ptr(), B, _p_i32(flat), _p_i32(lens), C.byref(c),
                                         _p_i32(toks), max_out, _p_i32(ntok), _p_f32(slp), _p_f32(nsp),
                                         self._stream()))
        del keep
        return [DecodeResult(toks[b, :ntok[b]].tolist(), float(slp[b]), float(nsp[b])) for b in range(B)]

    def decoder_logits(self, cross_kv: torch.Tensor, seqs: Sequence[Sequence[int]]) -> List[torch.Tensor]:
        """Teacher-forced logits (fp32 [len, n_vocab]) of every sequence; parity / debug hook."""
        B = len(seqs)
        flat = _i32([t for p in seqs for t in p])
        lens = _i32([len(p) for p in seqs])
        total = int(lens.sum())
        buf = torch.empty((total, self.ldv), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.b200w_decoder_logits(self._h, cross_kv.data_ptr(), B, _p_i32(flat), _p_i32(lens),
                                                 buf.data_ptr(), self._stream()))
        out, o = [], 0
        for n in lens:
            out.append(buf[o:o + n, : self.dims.n_vocab])
            o += int(n)
        return out

    def detect_language(self, cross_kv: torch.Tensor, sot: int, lang_ids: Sequence[int]) -> np.ndarray:
        B = cross_kv.shape[0]
        ids = _i32(list(lang_ids))
        probs = np.zeros((B, len(ids)), dtype=np.float32)
        _lib.check(self.lib.b200w_detect_language(self._h, cross_kv.data_ptr(), B, sot, _p_i32(ids), len(ids),
                                                  _p_f32(probs), self._stream()))
        return probs

    def align(self, cross_kv: torch.Tensor, seqs: Sequence[Sequence[int]], num_frames: Sequence[int],
              sot_len: int, heads: Sequence[Tuple[int, int]], median_width: int = 7):
        """Word-timing pass.  seqs[b] = sot_sequence + [no_timestamps] + text + [eot].
        Returns per sequence (text_indices, time_indices, text_token_probs)."""
        B = len(seqs)
        assert cross_kv.shape[0] == B and B <= self.max_batch
        flat = _i32([t for p in seqs for t in p])
        lens = _i32([len(p) for p in seqs])
        nf = _i32(list(num_frames))
        hd = _i32([x for lh in heads for x in lh])
        max_path = int(lens.max()) + self.dims.n_audio_ctx + 2
        max_tok = int(lens.max())
        path = np.zeros((B, max_path, 2), dtype=np.int32)
        plen = np.zeros(B, dtype=np.int32)
        tprob = np.zeros((B, max_tok), dtype=np.float32)
        _lib.check(self.lib.b200w_align(self._h, cross_kv.data_ptr(), B, _p_i32(flat), _p_i32(lens), _p_i32(nf),
                                        sot_len, _p_i32(hd), len(heads), median_width, _p_i32(path), max_path,
                                        _p_i32(plen), _p_f32(tprob), max_tok, self._stream()))
        out = []
        for b in range(B):
            n_text = int(lens[b]) - sot_len - 2
            out.append((path[b, :plen[b], 0].copy(), path[b, :plen[b], 1].copy(), tprob[b, :max(n_text, 0)].copy()))
        return out
