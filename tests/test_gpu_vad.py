"""the batched VAD gate returns, for every stream and every tick, exactly what the UNCHANGED reference
DynamicVADIterator returned on the same audio (tests/golden/vad_gate_golden.json, recorded by tools/make_vad_golden.py
from /root/reference/asr_components.py:12-78) -- including the live retuning of the silence threshold."""
import json
import os

import numpy as np
import pytest

from tests.vad_scenarios import N_STREAMS, chunks_of, retunes, stream_audio

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vad_gate_golden.json")


def test_batched_vad_gate_matches_reference_events(cuda_device):
    from realtime_whisper_asr_b200.vad import BatchedVADGate
    gold = json.load(open(GOLD))
    ids = list(range(N_STREAMS))
    gate = BatchedVADGate(ids, initial_silence_ms=500, min_silence_ms=200, max_silence_ms=1000, threshold=0.5,
                          energy_k=0.02)
    chunks = {s: chunks_of(stream_audio(s), s) for s in ids}
    n_ticks = max(len(c) for c in chunks.values())
    n_events = 0
    for tick in range(n_ticks):
        for s in ids:
            for t, ms in retunes(s):
                if t == tick:
                    gate.set_silence_duration(s, ms)
        feed = {s: chunks[s][tick] for s in ids if tick < len(chunks[s])}
        got = gate(feed)  # ONE scoring launch + ONE gate launch for all streams of the tick
        for s in feed:
            want = gold[str(s)][tick]
            assert got[s] == want, (s, tick, got[s], want)
            n_events += want is not None
    assert n_events >= 5 * N_STREAMS and gate.windows_scored > 1000
    # reset_states: the stream starts over (first voiced window reports a start at -pad again)
    gate.reset_states(0)
    assert gate({0: stream_audio(0)[:2048]})[0] == {"start": -1600}


def test_custom_scorer_and_empty_ticks(cuda_device):
    import torch
    from realtime_whisper_asr_b200.vad import BatchedVADGate
    calls = []

    def scorer(windows):  # e.g. a batched Silero module; here: voiced iff the window's mean is positive
        calls.append(windows.shape)
        return (windows.mean(dim=1) > 0).float()

    gate = BatchedVADGate(["a", "b"], scorer=scorer)
    x = np.ones(1024, np.float32)
    assert gate({"a": x[:100], "b": x[:200]}) == {"a": None, "b": None} and not calls   # no full window yet
    out = gate({"a": x, "b": -x})
    assert out["a"] == {"start": -1600} and out["b"] is None and calls[0][1] == 512
    out = gate({"a": -np.ones(16000, np.float32)})        # > 500 ms below threshold: the segment ends
    assert out["a"] is not None and "end" in out["a"]
    assert isinstance(gate.state, torch.Tensor)
