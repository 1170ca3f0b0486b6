"""``B200FilterSeparator``: the reference's band-pass "vocal separation" pre-filter on the GPU.

Mirrors ``SimpleFilterSeparator`` (/root/reference/vocal_separation.py:303-372; built by ``create_separator("filter")``
:375, enabled with ``method: "filter"`` in the shipped config.json, called per incoming chunk at
一键实时识别麦克风.py:1474-1482): ``separate(audio) -> (vocal_audio, background_audio)`` where ``vocal_audio =
scipy.signal.filtfilt(*butter(4, [85, 3400] Hz, 'band'), audio)`` and ``background = audio - vocal``.  ``separate_batch``
filters the chunks of many streams in one launch (one thread per chunk; see csrc/frontend.cu), which is what a
multi-stream scheduler calls once per tick; the filtered chunk stays on the device for the dedup features and the
log-mel kernel that read it next.  On any failure the reference returns the original audio with a warning; so does this.
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import frontend


class B200FilterSeparator:
    def __init__(self, sample_rate: int = 16000, low_cut: float = 85.0, high_cut: float = 3400.0, device=None):
        self.sample_rate = sample_rate
        self.low_cut = low_cut
        self.high_cut = high_cut
        self.enabled = True
        self.device = frontend.require_cuda(device)
        self.coeffs = frontend.bandpass_coefficients(low_cut, high_cut, sample_rate)

    def is_available(self) -> bool:
        return torch.cuda.is_available()

    def separate_device(self, audio_dev: torch.Tensor, offsets: np.ndarray) -> torch.Tensor:
        """Filter chunks that are already on the device; the result stays there."""
        return frontend.bandpass_filter(audio_dev, offsets, self.coeffs)

    def separate_batch(self, chunks: Sequence[np.ndarray]) -> List[Tuple[np.ndarray, Optional[np.ndarray]]]:
        mono = [np.mean(c, axis=1) if np.asarray(c).ndim > 1 else np.asarray(c) for c in chunks]
        try:
            dev, offsets = frontend.upload_audio([m.astype(np.float32) for m in mono], self.device)
            out = self.separate_device(dev, offsets).cpu().numpy()
        except Exception as e:  # the reference's convention: warn and hand the original audio back
            warnings.warn(f"band-pass filter failed: {e}; returning the original audio")
            return [(m, None) for m in mono]
        res = []
        for m, a, b in zip(mono, offsets[:-1], offsets[1:]):
            if len(m) <= frontend.BANDPASS_PAD:  # scipy.filtfilt refuses such inputs; the reference then keeps them
                res.append((m, None))
                continue
            vocal = out[a:b].astype(np.float64)   # the reference's filtfilt returns float64
            res.append((vocal, m - vocal))
        return res

    def separate(self, audio: np.ndarray) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        return self.separate_batch([audio])[0]
