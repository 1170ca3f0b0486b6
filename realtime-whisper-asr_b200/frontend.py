"""Host side of the fused front end: log-mel features and audio-dedup features on the GPU.

Replaces faster-whisper's ``FeatureExtractor.__call__`` (NumPy; reached from
``/root/reference/asr_components.py:279-288``) and the arithmetic of
``/root/reference/audio_deduplicator.py:60-181``.  PyTorch is used for device buffers only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from . import _lib

HOP = 160
SAMPLE_RATE = 16000
N_FRAMES = 3000
TILE_FRAMES = 32


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.B200WError("CUDA device required: the B200 backend has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


@dataclass
class MelBatch:
    """Log-mel features of a batch of clips, resident on the device.

    feats: flat fp32 tensor; clip s is ``feats[offsets[s]:offsets[s+1]].view(n_mels, n_frames[s])``.
    """
    feats: torch.Tensor
    offsets: List[int]
    n_frames: List[int]
    n_mels: int
    feat_offsets_dev: torch.Tensor
    n_frames_dev: torch.Tensor

    def clip(self, s: int) -> torch.Tensor:
        return self.feats[self.offsets[s]: self.offsets[s + 1]].view(self.n_mels, self.n_frames[s])


def upload_audio(audios: Sequence[np.ndarray], device) -> tuple[torch.Tensor, np.ndarray]:
    """Concatenate clips into one pinned host buffer and copy it to the device (async)."""
    lens = np.array([len(a) for a in audios], dtype=np.int64)
    offsets = np.zeros(len(audios) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    host = torch.empty(max(total, 1), dtype=torch.float32, pin_memory=True)
    hv = host.numpy()
    for a, o in zip(audios, offsets[:-1]):
        hv[o:o + len(a)] = np.asarray(a, dtype=np.float32)
    return host.to(device, non_blocking=True), offsets


def log_mel(audio_dev: torch.Tensor, offsets: np.ndarray, n_mels: int) -> MelBatch:
    """Batched log-mel of clips already on the device (offsets: host int64 [n+1])."""
    lib = _lib.load()
    device = audio_dev.device
    n = len(offsets) - 1
    lens = np.diff(offsets)
    n_frames = (lens // HOP + 1).astype(np.int64)
    out_sizes = n_frames * n_mels
    out_offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(out_sizes, out=out_offsets[1:])
    tiles = (n_frames + TILE_FRAMES - 1) // TILE_FRAMES
    tile_prefix = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(tiles, out=tile_prefix[1:])
    # one small metadata upload: [offsets | out_offsets] int64, [tile_prefix | n_frames] int32
    meta64 = torch.from_numpy(np.concatenate([offsets, out_offsets])).to(device, non_blocking=True)
    meta32 = torch.from_numpy(np.concatenate([tile_prefix, n_frames.astype(np.int32)])).to(
        device, non_blocking=True)
    feats = torch.empty(int(out_offsets[-1]), dtype=torch.float32, device=device)
    smax = torch.empty(n, dtype=torch.int32, device=device)
    off_dev = meta64[: n + 1]
    oo_dev = meta64[n + 1:]
    tp_dev = meta32[: n + 1]
    nf_dev = meta32[n + 1:]
    _lib.check(lib.b200w_logmel(audio_dev.data_ptr(), off_dev.data_ptr(), oo_dev.data_ptr(),
                                tp_dev.data_ptr(), n, int(tile_prefix[-1]), n_mels, feats.data_ptr(),
                                smax.data_ptr(), int(n_frames.max()), _stream_ptr(device)))
    return MelBatch(feats, [int(x) for x in out_offsets], [int(x) for x in n_frames], n_mels,
                    oo_dev, nf_dev)


def pack_windows(mel: MelBatch, win_stream: Sequence[int], seeks: Sequence[int]) -> torch.Tensor:
    """Encoder input windows bf16 [n_windows, 3000, 128] (zero-padded features, time-major)."""
    lib = _lib.load()
    device = mel.feats.device
    nw = len(win_stream)
    meta = torch.tensor(list(seeks) + list(win_stream), dtype=torch.int32).to(device, non_blocking=True)
    windows = torch.empty((nw, N_FRAMES, 128), dtype=torch.bfloat16, device=device)
    _lib.check(lib.b200w_pack_windows(mel.feats.data_ptr(), mel.feat_offsets_dev.data_ptr(),
                                      mel.n_frames_dev.data_ptr(), meta[:nw].data_ptr(),
                                      meta[nw:].data_ptr(), nw, mel.n_mels, windows.data_ptr(),
                                      _stream_ptr(device)))
    return windows


def _dedup_points(n: int) -> int:
    """Shared-memory footprint of a chunk in 8-byte units (mirrors dedup_features_kernel)."""
    m = n // 2 if n % 2 == 0 else n
    r = m
    for p in (2, 3, 5):
        while r > 1 and r % p == 0:
            r //= p
    p = 7
    smooth = True
    while r > 1:
        if r % p == 0:
            if p > 64:
                smooth = False
                break
            r //= p
        else:
            p += 2
            if p > 64 and r > 1:
                smooth = False
                break
    if smooth:
        return max(m, 1)
    return (4 * (n + n // 2 + 1) + 7) // 8  # direct DFT: samples + magnitudes as fp32


def dedup_features(audio_dev: torch.Tensor, offsets: np.ndarray, sample_rate: int = SAMPLE_RATE,
                   return_raw: bool = False):
    """AudioDeduplicator._extract_features for every chunk; returns fp32 [n_chunks, 5] on device."""
    lib = _lib.load()
    device = audio_dev.device
    n = len(offsets) - 1
    lens = np.diff(offsets)
    if n == 0:
        z = torch.zeros((0, 5), dtype=torch.float32, device=device)
        return (z, z) if return_raw else z
    max_points = max(_dedup_points(int(n_)) for n_ in set(lens.tolist()))
    if max_points > 27648:
        raise _lib.B200WError("dedup chunk too long for the shared-memory DFT "
                              "(N <= 55296 samples even, 27648 odd)")
    off_dev = torch.from_numpy(np.asarray(offsets, dtype=np.int64)).to(device, non_blocking=True)
    out = torch.empty((n, 5), dtype=torch.float32, device=device)
    raw = torch.empty((n, 5), dtype=torch.float32, device=device) if return_raw else None
    _lib.check(lib.b200w_dedup_features(audio_dev.data_ptr(), off_dev.data_ptr(), n, max_points,
                                        sample_rate, out.data_ptr(),
                                        raw.data_ptr() if raw is not None else None,
                                        _stream_ptr(device)))
    return (out, raw) if return_raw else out


# ----------------------------------------------------------------------------------- band-pass pre-filter
def butter_bandpass(order: int, low_hz: float, high_hz: float, sample_rate: int = SAMPLE_RATE):
    """Digital Butterworth band-pass as (b, a) float64 -- scipy.signal.butter(order, [low, high], btype='band') with
    normalised edges low/(sr/2), high/(sr/2): analog low-pass prototype poles, low-pass -> band-pass transform,
    bilinear transform (fs = 2, pre-warped edges), polynomial expansion.  2*order + 1 coefficients each."""
    nyq = sample_rate / 2.0
    wn = np.array([low_hz / nyq, high_hz / nyq], dtype=np.float64)
    if not (0.0 < wn[0] < wn[1] < 1.0):
        raise ValueError("band edges must satisfy 0 < low < high < sample_rate / 2")
    fs = 2.0
    warped = 2.0 * fs * np.tan(np.pi * wn / fs)
    bw, wo = warped[1] - warped[0], np.sqrt(warped[0] * warped[1])
    m = np.arange(-order + 1, order, 2)
    p_lp = -np.exp(1j * np.pi * m / (2 * order))            # prototype poles (no zeros, gain 1)
    p_s = p_lp * bw / 2.0
    p_bp = np.concatenate([p_s + np.sqrt(p_s ** 2 - wo ** 2), p_s - np.sqrt(p_s ** 2 - wo ** 2)])
    z_bp = np.zeros(order)                                  # `order` zeros at the origin
    k_bp = bw ** order
    fs2 = 2.0 * fs
    z_d = (fs2 + z_bp) / (fs2 - z_bp)
    p_d = (fs2 + p_bp) / (fs2 - p_bp)
    z_d = np.append(z_d, -np.ones(len(p_bp) - len(z_bp)))   # zeros at infinity map to z = -1
    k_d = k_bp * np.real(np.prod(fs2 - z_bp) / np.prod(fs2 - p_bp))
    b = k_d * np.real(np.poly(z_d))
    a = np.real(np.poly(p_d))
    return b.astype(np.float64), a.astype(np.float64)


def lfilter_zi(b: np.ndarray, a: np.ndarray) -> np.ndarray:
    """Steady-state initial conditions of the step response (scipy.signal.lfilter_zi) for len(a) == len(b)."""
    b, a = np.asarray(b, np.float64) / a[0], np.asarray(a, np.float64) / a[0]
    n = len(a)
    comp = np.zeros((n - 1, n - 1))
    comp[0, :] = -a[1:]
    comp[1:, :-1] = np.eye(n - 2)
    i_minus_a = np.eye(n - 1) - comp.T
    rhs = b[1:] - a[1:] * b[0]
    return np.linalg.solve(i_minus_a, rhs)


BANDPASS_PAD = 27


def bandpass_coefficients(low_hz: float = 85.0, high_hz: float = 3400.0, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """[b (9) | a (9) | zi (8)] float64 for the reference's 4th-order Butterworth band-pass
    (/root/reference/vocal_separation.py:351-357)."""
    b, a = butter_bandpass(4, low_hz, high_hz, sample_rate)
    return np.ascontiguousarray(np.concatenate([b, a, lfilter_zi(b, a)]))


def bandpass_filter(audio_dev: torch.Tensor, offsets: np.ndarray, coeffs: np.ndarray) -> torch.Tensor:
    """scipy.signal.filtfilt(b, a, chunk) for every chunk [offsets[i], offsets[i+1]) of a device buffer; returns a new
    fp32 device tensor of the same layout.  Chunks of <= 27 samples come back unchanged."""
    import ctypes as C
    lib = _lib.load()
    device = audio_dev.device
    n = len(offsets) - 1
    out = torch.empty_like(audio_dev)
    if n <= 0:
        return out
    assert coeffs.dtype == np.float64 and coeffs.shape == (26,)
    off_dev = torch.from_numpy(np.asarray(offsets, dtype=np.int64)).to(device, non_blocking=True)
    scratch = torch.empty((n, BANDPASS_PAD), dtype=torch.float32, device=device)
    _lib.check(lib.b200w_bandpass_filtfilt(audio_dev.data_ptr(), off_dev.data_ptr(), n,
                                           coeffs.ctypes.data_as(C.POINTER(C.c_double)), out.data_ptr(),
                                           scratch.data_ptr(), _stream_ptr(device)))
    return out
