"""``BatchedVADGate``: the reference's voice-activity gate for many streams at once.

The reference wraps ONE stream in ``DynamicVADIterator`` (/root/reference/asr_components.py:12-78): audio is re-chunked
to 512-sample windows, Silero scores each window on the CPU (``torch.hub`` model, one tiny inference per 32 ms per
stream), upstream's ``VADIterator`` state machine emits ``{'start': n}`` / ``{'end': n}`` sample events, and
``set_silence_duration`` retunes ``min_silence_samples`` live (clamped to [min, max] ms, ignored below a 50 ms change).
Here every stream's windows of a tick are scored in ONE launch and every stream's state machine advances in a second
one (csrc/frontend.cu: vad_window_energy_kernel, vad_gate_kernel); the per-stream state lives on the device.

The scorer is pluggable: ``scorer(windows_dev [W, 512] fp32) -> probs_dev [W] fp32``.  The built-in one is the energy
model ``rms / (rms + k)`` (what the tests use in place of the un-vendored Silero network, which cannot be fetched
offline); a batched Silero module can be passed instead -- its recurrent state is then the scorer's business.
"""
from __future__ import annotations

from typing import Callable, Dict, Hashable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, frontend

WINDOW = 512


class BatchedVADGate:
    def __init__(self, stream_ids: Sequence[Hashable], initial_silence_ms: float = 500, min_silence_ms: float = 200,
                 max_silence_ms: float = 1000, threshold: float = 0.5, sampling_rate: int = 16000,
                 speech_pad_ms: float = 100, scorer: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 energy_k: float = 0.02, device=None):
        self.device = frontend.require_cuda(device)
        self.lib = _lib.load()
        self.ids = list(stream_ids)
        self.index: Dict[Hashable, int] = {s: i for i, s in enumerate(self.ids)}
        n = len(self.ids)
        self.min_silence_ms, self.max_silence_ms = min_silence_ms, max_silence_ms
        self.current_silence_ms = [float(initial_silence_ms)] * n
        self.threshold = float(threshold)
        self.sampling_rate = sampling_rate
        self.speech_pad_samples = int(sampling_rate * speech_pad_ms / 1000)
        self.scorer = scorer
        self.energy_k = float(energy_k)
        self.state = torch.zeros((n, 3), dtype=torch.int64, device=self.device)  # triggered, temp_end, current_sample
        self._min_sil_host = np.full(n, sampling_rate * initial_silence_ms / 1000, dtype=np.float64)
        self.min_silence = torch.from_numpy(self._min_sil_host.copy()).to(self.device)
        self.leftover: List[np.ndarray] = [np.zeros(0, np.float32) for _ in range(n)]
        self.windows_scored = 0

    def set_silence_duration(self, sid: Hashable, silence_ms: float) -> bool:
        i = self.index[sid]
        silence_ms = max(self.min_silence_ms, min(self.max_silence_ms, silence_ms))
        if abs(silence_ms - self.current_silence_ms[i]) > 50:
            self.current_silence_ms[i] = silence_ms
            self._min_sil_host[i] = self.sampling_rate * silence_ms / 1000
            self.min_silence = torch.from_numpy(self._min_sil_host.copy()).to(self.device)
            return True
        return False

    def reset_states(self, sid: Hashable) -> None:
        i = self.index[sid]
        self.state[i].zero_()
        self.leftover[i] = np.zeros(0, np.float32)

    def __call__(self, chunks: Dict[Hashable, np.ndarray]) -> Dict[Hashable, Optional[dict]]:
        """One tick: ``chunks[sid]`` is that stream's new audio (any length).  Returns per stream what
        ``DynamicVADIterator.__call__`` would: ``None``, ``{'start': n}``, ``{'end': n}`` or both."""
        n = len(self.ids)
        counts = np.zeros(n + 1, dtype=np.int32)
        rows = []
        for sid, x in chunks.items():
            i = self.index[sid]
            buf = np.concatenate([self.leftover[i], np.asarray(x, dtype=np.float32).reshape(-1)])
            k = len(buf) // WINDOW
            counts[i + 1] = k
            if k:
                rows.append((i, buf[: k * WINDOW].reshape(k, WINDOW)))
            self.leftover[i] = buf[k * WINDOW:]
        offsets = np.cumsum(counts, dtype=np.int32)
        total = int(offsets[-1])
        out: Dict[Hashable, Optional[dict]] = {sid: None for sid in chunks}
        if total == 0:
            return out
        host = torch.empty((total, WINDOW), dtype=torch.float32, pin_memory=True)
        hv = host.numpy()
        for i, w in sorted(rows, key=lambda r: r[0]):
            hv[offsets[i]:offsets[i + 1]] = w
        windows = host.to(self.device, non_blocking=True)
        s = frontend._stream_ptr(self.device)
        if self.scorer is not None:
            probs = self.scorer(windows).to(torch.float32).contiguous()
        else:
            probs = torch.empty(total, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.b200w_vad_window_energy(windows.data_ptr(), total, self.energy_k, probs.data_ptr(), s))
        off_dev = torch.from_numpy(offsets).to(self.device, non_blocking=True)
        events = torch.empty((n, 4), dtype=torch.int64, device=self.device)
        _lib.check(self.lib.b200w_vad_gate(probs.data_ptr(), off_dev.data_ptr(), n, self.state.data_ptr(),
                                           self.min_silence.data_ptr(), self.threshold, self.speech_pad_samples,
                                           events.data_ptr(), s))
        self.windows_scored += total
        ev = events.cpu().numpy()
        for sid in chunks:
            i = self.index[sid]
            r = {}
            if ev[i, 0]:
                r["start"] = int(ev[i, 1])
            if ev[i, 2]:
                r["end"] = int(ev[i, 3])
            out[sid] = r or None
        return out
