"""Seeded synthetic 16 kHz inputs shared by tests and bench.py."""
import numpy as np

SR = 16000


def speech_like(seed: int, seconds: float) -> np.ndarray:
    """Harmonic stack with slow FM x syllabic AM + pink-ish noise at -30 dB, with silences."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * SR))
    if n == 0:
        return np.zeros(0, dtype=np.float32)
    t = np.arange(n) / SR
    f0 = rng.uniform(90, 250)
    fm = 1.0 + 0.05 * np.sin(2 * np.pi * rng.uniform(0.3, 1.0) * t + rng.uniform(0, 6.28))
    phase = 2 * np.pi * np.cumsum(f0 * fm) / SR
    x = np.zeros(n)
    for h in range(1, int(rng.integers(3, 6)) + 1):
        x += np.sin(h * phase + rng.uniform(0, 6.28)) / h
    am = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(3, 6) * t + rng.uniform(0, 6.28)))
    x *= am
    # silences of 0.3-1.0 s every 2-5 s
    pos = rng.uniform(0.5, 2.0)
    while pos < seconds:
        dur = rng.uniform(0.3, 1.0)
        a, b = int(pos * SR), min(n, int((pos + dur) * SR))
        x[a:b] = 0.0
        pos += dur + rng.uniform(2.0, 5.0)
    white = rng.standard_normal(n)
    pink = np.cumsum(white) * 0.02 + white  # crude 1/f tilt
    pink /= np.abs(pink).max() + 1e-9
    x = 0.3 * x / (np.abs(x).max() + 1e-9) + 10 ** (-30 / 20) * pink
    return np.clip(x, -1, 1).astype(np.float32)
