"""Oracle (test infrastructure): cross-attention DTW word timing on the CPU.

Restates CTranslate2 Whisper.align + faster-whisper find_alignment (third-party, un-vendored;
reached from /root/reference/asr_components.py:279-288 with word_timestamps=True; algorithm:
SURVEY.md Appendix A.5 == OpenAI whisper/timing.py).  PARITY UNPINNED by the reference.

CTranslate2 detail kept here: the attention probabilities come from the decoder layers as they
are (softmax over all 1500 encoder positions) and are cropped to num_frames//2 afterwards; they are
not re-normalised after the crop.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from .model import OracleWhisper


def median_filter(x: np.ndarray, width: int) -> np.ndarray:
    """Median along the last axis, reflect padding; skipped when the axis is too short."""
    pad = width // 2
    if x.shape[-1] <= pad or width <= 1:
        return x
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, width, axis=-1)
    return np.sort(win, axis=-1)[..., pad]


def dtw(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """OpenAI whisper dtw_cpu (same tie-breaking)."""
    N, M = x.shape
    cost = np.full((N + 1, M + 1), np.inf, dtype=np.float32)
    trace = -np.ones((N + 1, M + 1), dtype=np.int8)
    cost[0, 0] = 0
    for j in range(1, M + 1):
        for i in range(1, N + 1):
            c0, c1, c2 = cost[i - 1, j - 1], cost[i - 1, j], cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = x[i - 1, j - 1] + c
            trace[i, j] = t
    i, j = N, M
    trace[0, :] = 2
    trace[:, 0] = 1
    res = []
    while i > 0 or j > 0:
        res.append((i - 1, j - 1))
        t = trace[i, j]
        if t == 0:
            i -= 1
            j -= 1
        elif t == 1:
            i -= 1
        else:
            j -= 1
    res = np.array(res[::-1], dtype=np.int64).reshape(-1, 2)
    return res[:, 0], res[:, 1]


def align(model: OracleWhisper, enc_row: torch.Tensor, sot_sequence: Sequence[int], no_timestamps: int,
          eot: int, text_tokens: Sequence[int], num_frames: int, heads: Sequence[Tuple[int, int]],
          median_width: int = 7):
    """Returns (text_indices, time_indices, text_token_probs) like ctranslate2 align()."""
    tokens = list(sot_sequence) + [no_timestamps] + list(text_tokens) + [eot]
    logits, cw = model.teacher_forced_logits(tokens, enc_row, want_cross_weights=True)
    nf = max(1, min(model.dims.n_audio_ctx, num_frames // 2))
    w = np.stack([cw[l][0, h].numpy() for l, h in heads])[:, :, :nf]  # [heads, T, nf]
    mean = w.mean(axis=-2, keepdims=True, dtype=np.float32)
    std = np.sqrt(((w - mean) ** 2).mean(axis=-2, keepdims=True, dtype=np.float32))
    w = (w - mean) / std
    w = median_filter(w, median_width)
    matrix = w.mean(axis=0)
    matrix = matrix[len(sot_sequence):-1]
    ti, fi = dtw(-matrix)
    probs = torch.softmax(logits[len(sot_sequence):-2], dim=-1).numpy()
    tp = probs[np.arange(len(text_tokens)), list(text_tokens)] if len(text_tokens) else np.zeros(0, np.float32)
    return ti, fi, tp.astype(np.float32)
