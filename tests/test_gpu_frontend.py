"""GPU parity: fused log-mel front end and dedup features vs the CPU oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import frontend as OF
from realtime_whisper_asr_b200 import frontend as FE
from tests.synth import speech_like

pytestmark = pytest.mark.gpu

# log-mel tolerance (north_star: 1e-4 relative): |gpu - ref| <= 1e-4 * max(1, |ref|)
MEL_TOL = 1e-4


def _check_mel(audios, n_mels, device):
    audio_dev, offsets = FE.upload_audio(audios, device)
    mel = FE.log_mel(audio_dev, offsets, n_mels)
    torch.cuda.synchronize()
    worst = 0.0
    for s, a in enumerate(audios):
        ref = OF.log_mel(a, n_mels)
        got = mel.clip(s).cpu().numpy()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        worst = max(worst, float(err.max()))
    return worst


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_speech_like(cuda_device, n_mels):
    audios = [speech_like(1234 + i, seconds) for i, seconds in enumerate([30.0, 1.0, 7.3, 0.5, 12.345])]
    worst = _check_mel(audios, n_mels, cuda_device)
    assert worst <= MEL_TOL, worst


def test_logmel_edge_lengths(cuda_device):
    rng = np.random.default_rng(7)
    lens = [0, 1, 17, 159, 160, 161, 319, 400, 401, 5000, 480000]
    audios = [(rng.standard_normal(n) * 0.1).astype(np.float32) for n in lens]
    worst = _check_mel(audios, 80, cuda_device)
    assert worst <= MEL_TOL, worst


def test_logmel_edge_signals(cuda_device):
    n = 16000 * 3
    t = np.arange(n) / 16000.0
    audios = [
        np.zeros(n, dtype=np.float32),                                   # silence
        np.clip(3.0 * np.sin(2 * np.pi * 220 * t), -1, 1).astype(np.float32),  # clipped
        np.eye(1, n, n // 2, dtype=np.float32)[0],                       # single impulse
        (np.random.default_rng(3).standard_normal(n) * 0.3).astype(np.float32),  # noise
    ]
    # pure tones / impulses have bins at the 80 dB clamp edge where fp32 FFT rounding (in the
    # reference as much as here) decides the value: allow 1e-3 there
    worst = _check_mel(audios, 80, cuda_device)
    assert worst <= 1e-3, worst


def test_pack_windows(cuda_device):
    audios = [speech_like(1, 31.0), speech_like(2, 2.0)]
    audio_dev, offsets = FE.upload_audio(audios, cuda_device)
    mel = FE.log_mel(audio_dev, offsets, 80)
    wins = FE.pack_windows(mel, [0, 0, 1], [0, 3000, 0])
    torch.cuda.synchronize()
    for w, (s, seek) in enumerate([(0, 0), (0, 3000), (1, 0)]):
        ref = OF.pad_or_trim(OF.log_mel(audios[s], 80)[:, seek:seek + 3000])  # [80, 3000]
        got = wins[w].float().cpu().numpy()  # [3000, 128]
        assert np.all(got[:, 80:] == 0)
        assert np.abs(got[:, :80].T - ref).max() <= 1e-2  # bf16 rounding of values in [-1, 2]
        # zero padded FEATURES (not audio): columns past the clip are exactly 0
        nfr = len(audios[s]) // 160 + 1
        if nfr - seek < 3000:
            assert np.all(got[max(nfr - seek, 0):, :] == 0)


@pytest.mark.parametrize("n", [16000, 32000, 1600, 640, 6400, 17, 641, 2, 1, 12345, 16001 - 1])
def test_dedup_features(cuda_device, n):
    rng = np.random.default_rng(n)
    chunks = [speech_like(n + 5, n / 16000.0)[:n] if n >= 1600 else (rng.standard_normal(n) * 0.1).astype(np.float32),
              (rng.standard_normal(n) * 0.05).astype(np.float32)]
    chunks = [np.ascontiguousarray(c[:n]) for c in chunks]
    audio_dev, offsets = FE.upload_audio(chunks, cuda_device)
    got, raw = FE.dedup_features(audio_dev, offsets, return_raw=True)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    for i, c in enumerate(chunks):
        ref = OF.dedup_features(c)
        # roll-off is an index pick: near-ties may move it by one bin
        tol = np.array([2e-5, 2e-5, 2e-5, 2.0 / max(n, 1) + 2e-5, 2e-5]) * 4
        assert np.all(np.abs(got[i] - ref) <= tol + 1e-5 * np.abs(ref)), (n, got[i], ref)


def test_dedup_zero_and_empty(cuda_device):
    chunks = [np.zeros(1600, dtype=np.float32), np.zeros(0, dtype=np.float32), np.ones(800, dtype=np.float32) * 0.5]
    audio_dev, offsets = FE.upload_audio(chunks, cuda_device)
    got = FE.dedup_features(audio_dev, offsets).cpu().numpy()
    for i, c in enumerate(chunks):
        ref = OF.dedup_features(c).astype(np.float32)
        # a constant chunk has an exactly-zero spectrum off DC in pocketfft; here fp32 twiddle
        # rounding leaves ~1e-8 of leakage, which the bandwidth's f^2 weighting lifts to ~1e-4
        assert np.allclose(got[i], ref, atol=5e-4), (got[i], ref)
