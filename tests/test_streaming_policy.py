"""the batched streaming policy (realtime_whisper_asr_b200/streaming.py) commits exactly what the
UNCHANGED reference policy commits -- /root/reference/enhanced_asr_processor.py EnhancedOnlineASRProcessor with its
EnhancedHypothesisBuffer / DynamicBufferManager over upstream's OnlineASRProcessor -- on scripted hypothesis
sequences: same prompt sent to the backend, same (beg, end, text) returned, same buffer offset / length / trimming
threshold after every iteration, same finish().  The reference's outputs are recorded in
tests/golden/streaming_policy_golden.json (tools/make_streaming_golden.py, dev container); where the reference tree
is present the live classes are driven again next to the golden."""
import json
import os

import numpy as np
import pytest

from realtime_whisper_asr_b200.streaming import AudioBuffer, BatchedOnlineASR
from tests.streaming_scenarios import SCENARIOS, ScriptedModel, chunks_of

GOLD = os.path.join(os.path.dirname(__file__), "golden", "streaming_policy_golden.json")


class _Recording(ScriptedModel):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.prompts = []

    def transcribe_batch(self, audios, languages=None, initial_prompts=None, **kw):
        self.prompts.append(list(initial_prompts))
        assert kw["word_timestamps"] is True and kw["condition_on_previous_text"] is True and kw["beam_size"] == 5
        return super().transcribe_batch(audios, languages, initial_prompts, **kw)


def _close(a, b):
    if a is None or b is None:
        return a is b
    return abs(a - b) <= 1e-9


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_batched_policy_matches_reference_golden(name):
    with open(GOLD) as f:
        gold = json.load(f)[name]
    sc = SCENARIOS[name]
    model = _Recording(sc["seed"], fail_at=sc.get("fail_at"))
    pol = BatchedOnlineASR(model, agreement_n=sc["n"], buffer_trimming=("segment", sc["trim"]),
                           enable_dynamic_buffer=sc["dynamic"], logfile=open(os.devnull, "w"))
    pol.add_stream("s", "en")
    for it, (chunk, want) in enumerate(zip(chunks_of(sc), gold)):
        pol.insert_audio_chunk("s", chunk)
        n_before = len(model.prompts)
        got = pol.process_iter(["s"])["s"]
        prompt = model.prompts[-1][0] if len(model.prompts) > n_before else None
        st = pol.streams["s"]
        assert prompt == want["prompt"], (it, prompt, want["prompt"])
        assert _close(got[0], want["out"][0]) and _close(got[1], want["out"][1]) and got[2] == want["out"][2], (it, got, want["out"])
        assert _close(st.offset, want["offset"]) and len(st.audio) == want["buffer_len"], (it, st.offset, len(st.audio), want)
        assert st.trimming_sec == want["trimming_sec"], (it, st.trimming_sec, want["trimming_sec"])
    fin = pol.finish("s")
    want = gold[-1]
    assert _close(fin[0], want["finish"][0]) and _close(fin[1], want["finish"][1]) and fin[2] == want["finish"][2]
    assert _close(pol.streams["s"].offset, want["offset"])


def test_many_streams_one_backend_call_per_step():
    """All scenarios as concurrent streams of ONE policy object: one backend call per step serves every stream, and
    each stream still commits exactly its golden text."""
    with open(GOLD) as f:
        gold = json.load(f)
    names = [n for n in SCENARIOS if SCENARIOS[n]["n"] == 2 and SCENARIOS[n]["trim"] == 15 and "fail_at" not in SCENARIOS[n]
             and SCENARIOS[n]["dynamic"] is False] + ["n2_trim15"]
    model = _Recording(SCENARIOS["n2_trim15"]["seed"])
    pol = BatchedOnlineASR(model, agreement_n=2, buffer_trimming=("segment", 15), enable_dynamic_buffer=False)
    ids = [f"s{i}" for i in range(6)]
    for sid in ids:
        pol.add_stream(sid, "en")
    chunks = chunks_of(SCENARIOS["n2_trim15"])
    for it, chunk in enumerate(chunks):
        for sid in ids:
            pol.insert_audio_chunk(sid, chunk)
        out = pol.process_iter()
        for sid in ids:
            assert out[sid][2] == gold["n2_trim15"][it]["out"][2]
    assert pol.calls == len(chunks) and model.calls == len(chunks)   # one call per step, not one per stream
    assert pol.audio_seconds_submitted > 6 * 40  # every re-submitted buffer second counts (the product's RTFx)
    assert names


def test_audio_buffer_appends_and_trims_without_quadratic_copies():
    b = AudioBuffer(capacity=1000)
    ref = np.zeros(0, np.float32)
    rng = np.random.default_rng(0)
    ptr0 = None
    for i in range(400):
        a = rng.standard_normal(int(rng.integers(1, 700))).astype(np.float32)
        b.append(a)
        ref = np.append(ref, a)
        if i % 7 == 3:
            k = int(rng.integers(0, len(ref)))
            b.drop_front(k)
            ref = ref[k:]
        assert len(b) == len(ref) and np.array_equal(b.view(), ref)
        if i == 300:
            ptr0 = b._buf.ctypes.data
    assert b._buf.ctypes.data == ptr0  # steady state: no re-allocation, only in-place compaction
    b.clear()
    assert len(b) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present (GPU box)")
def test_golden_is_what_the_reference_does_now():
    """Dev container only: re-run the unchanged reference classes and compare with the committed golden."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    before = open(GOLD).read()
    subprocess.run([sys.executable, os.path.join(root, "tools", "make_streaming_golden.py")], check=True,
                   stdout=subprocess.DEVNULL)
    assert open(GOLD).read() == before
