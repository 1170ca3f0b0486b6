"""Seeded streams for the VAD-gate parity test (shared by tools/make_vad_golden.py and tests/test_gpu_vad.py)."""
import numpy as np

from tests.synth import speech_like

N_STREAMS = 12
SR = 16000


def stream_audio(sid: int) -> np.ndarray:
    """Speech-like bursts separated by true silences of 0.2 .. 1.5 s (so that start / end events and the
    min-silence rule both fire); amplitudes far from the energy scorer's threshold."""
    rng = np.random.default_rng(900 + sid)
    parts = []
    for k in range(6):
        parts.append(speech_like(1000 * sid + k, float(rng.uniform(0.6, 2.5))) * 3.0)
        parts.append(np.zeros(int(rng.uniform(0.2, 1.5) * SR), np.float32))
    return np.clip(np.concatenate(parts), -1, 1).astype(np.float32)


def chunks_of(x: np.ndarray, sid: int):
    """The reference reads 640 samples (0.04 s) at a time; odd streams get irregular chunk sizes."""
    if sid % 2 == 0:
        return [x[i:i + 640] for i in range(0, len(x), 640)]
    rng = np.random.default_rng(sid)
    out, i = [], 0
    while i < len(x):
        n = int(rng.integers(100, 2000))
        out.append(x[i:i + n])
        i += n
    return out


def retunes(sid: int):
    """(tick, silence_ms) calls to set_silence_duration, as the reference's interval heuristic would make."""
    return [(40, 250), (120, 900), (200, 520)] if sid % 3 == 0 else []
