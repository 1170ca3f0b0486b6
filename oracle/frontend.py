"""Oracle (test infrastructure): log-mel front end and audio-dedup features on the CPU.

log_mel         restates faster-whisper FeatureExtractor.__call__ (third-party,
                un-vendored; call site /root/reference/asr_components.py:279-288;
                algorithm: SURVEY.md Appendix A.3).  PARITY UNPINNED by the reference.
dedup_features  restates /root/reference/audio_deduplicator.py:60-181 line by line.
                PINNED by tests/golden/dedup_golden.npz (outputs of the reference class).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_FREQ = N_FFT // 2 + 1  # 201


def _hz_to_mel(f):
    # slaney scale (librosa.hz_to_mel, htk=False)
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filters(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular filterbank [n_mels, n_fft//2+1] float32.

    Same construction as faster-whisper's FeatureExtractor.get_mel_filters (a NumPy
    restatement of librosa.filters.mel), SURVEY.md A.3 step 5.
    """
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2)
    mel_f = _mel_to_hz(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f.reshape(-1, 1) - fftfreqs.reshape(1, -1)
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def stft_power(audio: np.ndarray, padding: int = HOP) -> np.ndarray:
    """|STFT|^2 with the trailing frame dropped: [201, n_frames] float32 (A.3 steps 1-4)."""
    x = np.asarray(audio)
    if x.dtype != np.float32:
        x = x.astype(np.float32)
    if padding:
        x = np.pad(x, (0, padding))
    window = np.hanning(N_FFT + 1)[:-1].astype(np.float32)
    x = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + (x.shape[-1] - N_FFT) // HOP
    frames = np.lib.stride_tricks.as_strided(
        x, shape=(n_frames, N_FFT), strides=(x.strides[-1] * HOP, x.strides[-1]), writeable=False
    )
    frames = frames * window
    stft = np.fft.rfft(frames, n=N_FFT, axis=-1).astype(np.complex64).T  # [201, n_frames]
    return np.abs(stft[..., :-1]) ** 2


def log_mel(audio: np.ndarray, n_mels: int = 80, padding: int = HOP) -> np.ndarray:
    """faster-whisper (>=1.1) log-mel: [n_mels, len(audio)//160 + 1] float32.

    NB the clamp floor uses the max over the WHOLE call (all frames, all bins), and
    the result is NOT yet padded/cropped to 3000 frames (that is A.2 step 6).
    """
    mag = stft_power(audio, padding)
    mel_spec = mel_filters(n_mels) @ mag
    log_spec = np.log10(np.clip(mel_spec, a_min=1e-10, a_max=None))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec.astype(np.float32)


def pad_or_trim(features: np.ndarray, length: int = 3000) -> np.ndarray:
    """faster-whisper pad_or_trim on the LAST axis: zero-pads the FEATURES (A.2 step 6)."""
    n = features.shape[-1]
    if n > length:
        return features[..., :length]
    if n < length:
        pad = [(0, 0)] * (features.ndim - 1) + [(0, length - n)]
        return np.pad(features, pad)
    return features


# --------------------------------------------------------------------------------------
# audio-dedup features: /root/reference/audio_deduplicator.py:60-181
# --------------------------------------------------------------------------------------
def dedup_features_raw(audio: np.ndarray, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """The five features BEFORE the max-abs normalisation (float64), :92-155."""
    if len(audio) == 0:
        return np.zeros(5)
    rms = np.sqrt(np.mean(audio ** 2))  # :92
    fft = np.fft.rfft(audio)  # :99
    magnitude = np.abs(fft)
    frequencies = np.fft.rfftfreq(len(audio), 1.0 / sample_rate)  # :103
    centroid = np.sum(frequencies * magnitude) / (np.sum(magnitude) + 1e-10)  # :105
    centroid = centroid / (sample_rate / 2)  # :107
    if len(audio) > 1:  # :115-118
        zcr = np.sum(np.diff(np.signbit(audio))) / len(audio)
    else:
        zcr = 0.0
    cumsum = np.cumsum(magnitude)  # :128
    total = cumsum[-1]
    if total > 1e-10:
        idx = np.where(cumsum >= 0.85 * total)[0]  # :131-132
        rolloff = frequencies[idx[0]] / (sample_rate / 2) if len(idx) > 0 else 1.0
    else:
        rolloff = 0.0
    if centroid > 0:  # :148-153
        cf = centroid * sample_rate / 2
        bw = np.sqrt(np.sum(((frequencies - cf) ** 2) * magnitude) / (np.sum(magnitude) + 1e-10))
        bw = bw / (sample_rate / 2)
    else:
        bw = 0.0
    return np.array([rms, centroid, zcr, rolloff, bw], dtype=np.float64)


def dedup_features(audio: np.ndarray, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """AudioDeduplicator._extract_features: five features / max|f| as float32, :161-176."""
    if len(audio) == 0:
        return np.zeros(5)
    f = dedup_features_raw(audio, sample_rate).astype(np.float32)
    m = np.max(np.abs(f))
    if m > 1e-10:
        return f / m
    return np.zeros(5, dtype=np.float32)


def cosine_similarity01(a: np.ndarray, b: np.ndarray) -> float:
    """AudioDeduplicator._cosine_similarity, :183-207."""
    n1, n2 = np.linalg.norm(a), np.linalg.norm(b)
    if n1 < 1e-10 or n2 < 1e-10:
        return 0.0
    return float((np.dot(a, b) / (n1 * n2) + 1.0) / 2.0)
