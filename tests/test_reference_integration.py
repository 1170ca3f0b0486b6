"""CPU: the UNCHANGED reference files on top of this backend's Python surface (dev container only:
/root/reference is not shipped to the GPU box).  A stub engine stands in for the CUDA model so the
test exercises signatures, kwarg forwarding and the streaming policy, not arithmetic."""
import inspect
import os
import sys

import numpy as np
import pytest

from tests import harness

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


class _Word:
    def __init__(self, s, e, w):
        self.start, self.end, self.word, self.probability = s, e, w, 0.9


class _Seg:
    def __init__(self, words, nsp=0.0):
        self.words, self.no_speech_prob = words, nsp
        self.start, self.end = words[0].start, words[-1].end
        self.text = "".join(w.word for w in words)


class StubModel:
    """WhisperModel-shaped: one word per started second of audio, deterministic."""
    calls = []

    def __init__(self, *a, **kw):
        StubModel.calls.append(("init", a, kw))

    def transcribe(self, audio, **kw):
        StubModel.calls.append(("transcribe", len(audio), kw))
        n = int(len(audio) // 16000)
        segs = []
        for s0 in range(0, n, 3):
            words = [_Word(float(i), i + 0.9, f" w{i}") for i in range(s0, min(n, s0 + 3))]
            segs.append(_Seg(words))
        return iter(segs), None


@pytest.fixture()
def ref_modules(monkeypatch):
    assert harness.install(REF)
    import types
    fw = types.ModuleType("faster_whisper")
    fw.WhisperModel = StubModel
    monkeypatch.setitem(sys.modules, "faster_whisper", fw)
    StubModel.calls.clear()
    import asr_components
    return asr_components


def test_constructor_signature_matches_reference(ref_modules):
    from realtime_whisper_asr_b200.backend import B200WhisperASR
    from whisper_online import FasterWhisperASR
    Custom = ref_modules.create_custom_faster_whisper_asr(FasterWhisperASR)
    ref_params = list(inspect.signature(Custom.__init__).parameters)
    ours = list(inspect.signature(B200WhisperASR.__init__).parameters)
    assert ours[: len(ref_params)] == ref_params
    ref_defaults = {k: v.default for k, v in inspect.signature(Custom.__init__).parameters.items()}
    our_defaults = {k: v.default for k, v in inspect.signature(B200WhisperASR.__init__).parameters.items()}
    for k in ref_params[1:]:
        assert our_defaults[k] == ref_defaults[k] or k == "logfile", k
    for name in ("transcribe", "ts_words", "segments_end_ts", "set_translate_task", "use_vad", "load_model"):
        assert callable(getattr(B200WhisperASR, name)) and callable(getattr(Custom, name))
    assert B200WhisperASR.sep == Custom.sep == ""


def test_b2_insertion_reference_factory_loads_our_model(ref_modules):
    """create_custom_faster_whisper_asr (unchanged) builds whatever faster_whisper.WhisperModel is."""
    from whisper_online import FasterWhisperASR
    Custom = ref_modules.create_custom_faster_whisper_asr(FasterWhisperASR)
    asr = Custom("auto", "large-v3", cache_dir="/tmp/x", device="cuda", compute_type="float16", device_index=0,
                 num_workers=2, transcribe_kwargs={"beam_size": 1, "temperature": 0.0})
    kind, args, kw = StubModel.calls[0]
    assert kind == "init" and args == ("large-v3",)
    assert kw == {"device": "cuda", "compute_type": "float16", "download_root": "/tmp/x", "num_workers": 2,
                  "device_index": 0}
    import inspect as _i
    from realtime_whisper_asr_b200.transcribe import B200WhisperModel
    accepted = set(_i.signature(B200WhisperModel.__init__).parameters)
    assert set(kw) <= accepted  # our engine accepts exactly what the reference forwards
    res = asr.transcribe(np.zeros(16000 * 4, np.float32), init_prompt="hi")
    kind, n, tkw = StubModel.calls[-1]
    assert set(tkw) == {"language", "initial_prompt", "beam_size", "temperature", "word_timestamps",
                        "condition_on_previous_text"}
    accepted_t = set(_i.signature(B200WhisperModel.transcribe_batch).parameters) | {"language"}
    assert set(tkw) <= accepted_t
    assert tkw["language"] is None and tkw["beam_size"] == 1 and tkw["word_timestamps"] is True
    assert asr.ts_words(res)[0] == (0.0, 0.9, " w0") and asr.segments_end_ts(res) == [2.9, 3.9]


def test_reference_streaming_policy_runs_on_our_backend_class(ref_modules, monkeypatch):
    """enhanced_asr_processor.py (unchanged) drives B200WhisperASR (stub engine injected)."""
    from realtime_whisper_asr_b200.backend import B200WhisperASR
    real_exists = os.path.exists
    monkeypatch.setattr(os.path, "exists", lambda p: True if "whisper_streaming-main" in str(p) else real_exists(p))
    import importlib
    eap = importlib.import_module("enhanced_asr_processor")
    asr = B200WhisperASR("en", model=StubModel(), transcribe_kwargs={"beam_size": 5})
    online = eap.EnhancedOnlineASRProcessor(asr, tokenizer=None, buffer_trimming=("segment", 15))
    committed = []
    for sec in range(20):
        online.insert_audio_chunk(np.zeros(16000, np.float32))
        beg, end, text = online.process_iter()
        if text:
            committed.append((beg, end, text))
    beg, end, text = online.finish()
    assert committed, "LocalAgreement committed nothing"
    words = "".join(t for _, _, t in committed).split()
    # (the stub names words by their second inside the CURRENT buffer, so names restart after a trim)
    assert words[:3] == ["w0", "w1", "w2"] and len(words) >= 10
    assert len(online.audio_buffer) / 16000 <= 16.5  # trimmed at segment boundaries past 15 s
    kinds = [c for c in StubModel.calls if c[0] == "transcribe"]
    assert len(kinds) == 20 and kinds[-1][2]["initial_prompt"] is not None


def test_dedup_class_has_reference_surface():
    assert harness.install(REF)
    from audio_deduplicator import AudioDeduplicator
    from realtime_whisper_asr_b200.dedup import B200AudioDeduplicator
    for name in ("should_skip", "get_stats", "reset_stats", "reset", "_extract_features", "_cosine_similarity"):
        assert hasattr(B200AudioDeduplicator, name) and hasattr(AudioDeduplicator, name)
    a = list(inspect.signature(AudioDeduplicator.__init__).parameters)
    b = list(inspect.signature(B200AudioDeduplicator.__init__).parameters)
    assert b[: len(a)] == a
    assert list(inspect.signature(AudioDeduplicator.should_skip).parameters) == \
        list(inspect.signature(B200AudioDeduplicator.should_skip).parameters)
