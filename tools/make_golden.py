"""Generate tests/golden/*.npz in the dev container (where /root/reference is mounted).

dedup_golden.npz  outputs of the REFERENCE class itself
                  (/root/reference/audio_deduplicator.py AudioDeduplicator._extract_features and
                  should_skip) on seeded synthetic chunks -> pins oracle.frontend.dedup_features.
logmel_golden.npz outputs of oracle.frontend.log_mel cross-checked here against
                  torch.stft + transformers.audio_utils.mel_filter_bank (the reference ships no
                  log-mel vectors: faster-whisper is not installable offline) -> regression pin.
Run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from audio_deduplicator import AudioDeduplicator  # noqa: E402  (the reference)

from oracle import frontend as OF  # noqa: E402
from tests.synth import speech_like  # noqa: E402


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    d = AudioDeduplicator()
    rng = np.random.default_rng(2024)
    lens = [16000, 32000, 1600, 640, 6400, 12345, 641, 17, 2, 1]
    chunks, feats = {}, {}
    for i, n in enumerate(lens):
        x = speech_like(500 + i, n / 16000.0)[:n] if n >= 1600 else (rng.standard_normal(n) * 0.1).astype(np.float32)
        x = np.ascontiguousarray(x[:n])
        chunks[f"chunk_{i}"] = x
        feats[f"feat_{i}"] = np.asarray(d._extract_features(x), dtype=np.float32)
    # should_skip decisions over a stream with repeats (reference history / threshold logic)
    d2 = AudioDeduplicator()
    base = [speech_like(900 + i, 1.0) for i in range(4)]
    seq = [base[0], base[1], base[0], base[2], base[2] * 0.999, base[3], base[1]]
    decisions, sims = [], []
    for k, a in enumerate(seq):
        skip, reason, details = d2.should_skip(a, 16000, current_time=100.0 + 0.2 * k)
        decisions.append(int(skip))
        sims.append(details["similarity"] if details else -1.0)
    np.savez_compressed(os.path.join(out, "dedup_golden.npz"), lens=np.array(lens), decisions=np.array(decisions),
                        sims=np.array(sims, dtype=np.float64), **chunks, **feats)

    import torch
    from transformers.audio_utils import mel_filter_bank
    mel_out = {}
    for j, (seed, sec, nm) in enumerate([(11, 2.0, 80), (12, 0.37, 128), (13, 5.5, 128)]):
        x = speech_like(seed, sec)
        m = OF.log_mel(x, nm)
        xp = np.pad(x, (0, 160))
        st = torch.stft(torch.from_numpy(xp), 400, 160, window=torch.hann_window(400), return_complex=True)
        mag = (st[:, :-1].abs() ** 2).numpy()
        fb = mel_filter_bank(201, nm, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T.astype(np.float32)
        ls = np.log10(np.clip(fb @ mag, 1e-10, None))
        ls = (np.maximum(ls, ls.max() - 8.0) + 4.0) / 4.0
        assert np.abs(ls - m).max() < 2e-4, np.abs(ls - m).max()
        mel_out[f"audio_{j}"] = x
        mel_out[f"mel_{j}"] = m
        mel_out[f"nmels_{j}"] = np.array(nm)
    np.savez_compressed(os.path.join(out, "logmel_golden.npz"), **mel_out)
    print("wrote", os.listdir(out))


if __name__ == "__main__":
    main()
