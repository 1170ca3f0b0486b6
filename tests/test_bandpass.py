"""the band-pass pre-filter.  CPU: the coefficient design and initial conditions equal scipy's.
GPU: the filtered chunks equal scipy.signal.filtfilt -- the call the reference makes
(/root/reference/vocal_separation.py:351-357) -- within 1e-5."""
import numpy as np
import pytest
from scipy import signal

from realtime_whisper_asr_b200 import frontend as FE
from tests.synth import speech_like


def test_butterworth_design_equals_scipy():
    for lo, hi, sr in ((85.0, 3400.0, 16000), (300.0, 3000.0, 16000), (85.0, 3400.0, 44100)):
        b, a = FE.butter_bandpass(4, lo, hi, sr)
        rb, ra = signal.butter(4, [lo / (sr / 2), hi / (sr / 2)], btype="band")
        assert np.allclose(b, rb, rtol=1e-9, atol=1e-15) and np.allclose(a, ra, rtol=1e-9, atol=1e-15)
        # (an ill-conditioned 8 x 8 solve: scipy's own answer moves by ~1e-7 with the solver)
        assert np.abs(FE.lfilter_zi(b, a) - signal.lfilter_zi(rb, ra)).max() <= 1e-6
    c = FE.bandpass_coefficients()
    assert c.shape == (26,) and abs(c[9] - 1.0) < 1e-15


@pytest.mark.gpu
def test_bandpass_matches_scipy_filtfilt(cuda_device):
    from realtime_whisper_asr_b200.separation import B200FilterSeparator
    rb, ra = signal.butter(4, [85.0 / 8000, 3400.0 / 8000], btype="band")
    sep = B200FilterSeparator()
    rng = np.random.default_rng(0)
    chunks = [speech_like(50 + i, 0.04) for i in range(130)]                   # VAC reads: 640 samples, > 1 CTA
    chunks += [speech_like(7, 1.0), speech_like(8, 2.0), rng.standard_normal(28).astype(np.float32),
               rng.standard_normal(61).astype(np.float32), speech_like(9, 30.0), np.ones(1000, np.float32),
               (rng.standard_normal(4000) * 0.9).astype(np.float32)]
    got = sep.separate_batch(chunks)
    worst = 0.0
    for x, (vocal, bg) in zip(chunks, got):
        ref = signal.filtfilt(rb, ra, x)
        err = np.abs(vocal - ref).max() / max(1.0, np.abs(ref).max())
        worst = max(worst, err)
        assert vocal.shape == x.shape and np.allclose(bg, x - vocal)
    print(f"band-pass vs scipy.signal.filtfilt: worst error {worst:.2e} over {len(chunks)} chunks")
    assert worst <= 1e-5
    # short chunks (scipy raises, the reference keeps the original) and the single-chunk API
    short = rng.standard_normal(27).astype(np.float32)
    v, b = sep.separate(short)
    assert np.array_equal(v, short) and b is None
    v, b = sep.separate(chunks[130])
    assert np.abs(v - signal.filtfilt(rb, ra, chunks[130])).max() <= 1e-5
    # stereo input is averaged to mono first, as the reference does
    st = np.stack([chunks[131], chunks[131] * 0.5], axis=1)
    v, _ = sep.separate(st)
    assert np.abs(v - signal.filtfilt(rb, ra, st.mean(axis=1))).max() <= 1e-5
