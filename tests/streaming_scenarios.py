"""Scripted streaming scenarios shared by tools/make_streaming_golden.py (drives the unchanged reference policy) and
tests/test_streaming_policy.py (drives the batched policy).  Test infrastructure.

ScriptedModel plays the backend: the audio it receives encodes its absolute start time (sample value = absolute
sample index * 1e-7), the "true" transcript is word k spoken during [0.37 k, 0.37 k + 0.3) s, grouped into segments
of 6 words; a call returns the words that start inside the buffer, the last few of them still unstable (text and
end time change from call to call, deterministically in (seed, absolute time, buffer length)) -- the behaviour
LocalAgreement exists for."""
from collections import namedtuple

import numpy as np

Word = namedtuple("Word", "start end word probability")
Segment = namedtuple("Segment", "start end words no_speech_prob")
SR = 16000
STEP, DUR, PER_SEG = 0.37, 0.3, 6
_VOCAB = "alpha bravo charlie delta echo foxtrot golf hotel india juliet kilo lima mike november oscar papa".split()

SCENARIOS = {
    "n2_trim15": dict(seed=1, n=2, trim=15, dynamic=False, seconds=40, chunk=1.0),
    "n2_dynamic": dict(seed=2, n=2, trim=15, dynamic=True, seconds=60, chunk=1.0),
    "n3_trim10": dict(seed=3, n=3, trim=10, dynamic=False, seconds=36, chunk=0.5),
    "n4_dynamic_small_chunks": dict(seed=4, n=4, trim=8, dynamic=True, seconds=30, chunk=0.25),
    "n2_backend_error": dict(seed=5, n=2, trim=15, dynamic=True, seconds=30, chunk=1.0, fail_at=12),
    "n2_irregular_chunks": dict(seed=6, n=2, trim=12, dynamic=True, seconds=45, chunk=None),
}


def chunks_of(sc):
    """Audio chunks of a scenario; sample i of the stream has the value i * 1e-7 (its absolute position)."""
    total = int(sc["seconds"] * SR)
    x = (np.arange(total, dtype=np.float64) * 1e-7).astype(np.float32)
    if sc["chunk"] is not None:
        n = int(sc["chunk"] * SR)
        return [x[i:i + n] for i in range(0, total, n)]
    rng = np.random.default_rng(sc["seed"])
    out, i = [], 0
    while i < total:
        n = int(rng.uniform(0.1, 2.5) * SR)
        out.append(x[i:i + n])
        i += n
    return out


class ScriptedModel:
    def __init__(self, seed, fail_at=None):
        self.seed, self.fail_at, self.calls = seed, fail_at, 0

    def _one(self, audio):
        audio = np.asarray(audio)
        if len(audio) == 0:
            return []
        t0 = round(float(audio[0]) * 1e7) / SR          # absolute start of the buffer
        length = len(audio) / SR
        k0 = int(np.ceil((t0 - 1e-9) / STEP))
        words = []
        k = k0
        while k * STEP < t0 + length - 0.05:
            words.append(k)
            k += 1
        rng = np.random.default_rng((self.seed, int(round(t0 * 1000)), len(audio)))
        segs, cur = [], []
        for i, k in enumerate(words):
            unstable = len(words) - i <= 3 and rng.random() < 0.5
            text = " " + _VOCAB[(k * 7 + self.seed) % len(_VOCAB)] + (str(k % 10) if k % 5 == 0 else "")
            if unstable:
                text = " " + _VOCAB[int(rng.integers(len(_VOCAB)))]
            start = k * STEP - t0
            end = start + DUR * (1.0 if not unstable else float(rng.uniform(0.5, 1.5)))
            cur.append(Word(round(start, 2), round(end, 2), text, 0.9))
            if k % PER_SEG == PER_SEG - 1:
                segs.append(Segment(cur[0].start, cur[-1].end, cur, 0.01))
                cur = []
        if cur:
            segs.append(Segment(cur[0].start, cur[-1].end, cur, 0.01))
        return segs

    def transcribe_batch(self, audios, languages=None, initial_prompts=None, **kw):
        self.calls += 1
        if self.fail_at is not None and self.calls == self.fail_at:
            raise RuntimeError("scripted backend failure")
        return [(self._one(a), None) for a in audios]
