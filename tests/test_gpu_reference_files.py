"""GPU: the UNCHANGED reference files driving the real CUDA backend.

The reference tree is read-only, is not part of this repository and does not exist on the GPU box, so this test
runs where BOTH a GPU and a reference tree are present: it looks for the tree at $B200W_REFERENCE_DIR, then
/root/reference, and is skipped otherwise.  (`tools/gpu_reference_check.sh` shows how it was run for
profiles/r02_reference_files_on_gpu.txt: a transient copy of the reference's .py files in a git-ignored directory
travels with one gpurun call and is deleted afterwards; nothing of the reference is committed.)

What it covers, with no edit to the reference:
  * B2: `create_custom_faster_whisper_asr(FasterWhisperASR)` (asr_components.py:182) loads B200WhisperModel through
    its own `load_model` (:232-265) once `install_as_faster_whisper()` ran, and `transcribe / ts_words /
    segments_end_ts` (:267-301) return words from the CUDA engine;
  * `EnhancedOnlineASRProcessor` (enhanced_asr_processor.py:239-398) commits text from 1 s increments on top of it,
    and commits exactly what `BatchedOnlineASR` commits on the same audio with the same backend;
  * `AudioDeduplicator` (audio_deduplicator.py) and `B200AudioDeduplicator` take the same decisions;
  * `DynamicVADIterator` (asr_components.py:12-78) and `BatchedVADGate` emit the same events.
"""
import os
import sys

import numpy as np
import pytest

from tests import harness
from tests.synth import speech_like

pytestmark = pytest.mark.gpu
_CANDIDATES = [os.environ.get("B200W_REFERENCE_DIR"), "/root/reference",
               os.path.join(os.path.dirname(os.path.abspath(__file__)), "_reference_copy")]
REF = next((p for p in _CANDIDATES if p and os.path.isfile(os.path.join(p, "asr_components.py"))), None)


@pytest.fixture(scope="module")
def ref(cuda_device):
    if REF is None:
        pytest.skip("no reference tree next to a GPU (see the module docstring)")
    assert harness.install(REF)
    os.environ["B200W_SYNTHETIC_WEIGHTS"] = "1"   # a size name: seeded synthetic weights (no checkpoint offline)
    from realtime_whisper_asr_b200.backend import install_as_faster_whisper
    install_as_faster_whisper()
    real_exists = os.path.exists
    os.path.exists = lambda p: True if "whisper_streaming-main" in str(p) else real_exists(p)
    try:
        import asr_components
        import enhanced_asr_processor
    finally:
        os.path.exists = real_exists
    yield asr_components, enhanced_asr_processor
    os.environ.pop("B200W_SYNTHETIC_WEIGHTS", None)


def test_unchanged_reference_backend_and_streaming_policy_on_the_gpu(ref):
    asr_components, eap = ref
    from whisper_online import FasterWhisperASR
    from realtime_whisper_asr_b200.streaming import BatchedOnlineASR
    from realtime_whisper_asr_b200.transcribe import B200WhisperModel
    Custom = asr_components.create_custom_faster_whisper_asr(FasterWhisperASR)
    asr = Custom("en", "tiny", cache_dir=None, device="cuda", compute_type="float16", device_index=0, num_workers=1,
                 transcribe_kwargs={"beam_size": 5, "temperature": 0.0})
    assert isinstance(asr.model, B200WhisperModel) and asr.model.synthetic and asr.sep == ""
    audio = speech_like(31, 24.0)
    res = asr.transcribe(audio[: 16000 * 6], init_prompt="previous words")
    words, ends = asr.ts_words(res), asr.segments_end_ts(res)
    assert res and words and len(ends) == len(res) and all(b >= a >= 0 for a, b, _ in words)
    # the reference's streaming policy, unchanged, on the CUDA backend -- next to the batched policy
    online = eap.EnhancedOnlineASRProcessor(asr, tokenizer=None, buffer_trimming=("segment", 15), agreement_n=2,
                                            logfile=open(os.devnull, "w"))
    pol = BatchedOnlineASR(asr.model, agreement_n=2, buffer_trimming=("segment", 15),
                           transcribe_kwargs={"beam_size": 5, "temperature": 0.0})
    pol.add_stream(0, "en")
    committed = []
    for i in range(24):
        chunk = audio[i * 16000:(i + 1) * 16000]
        online.insert_audio_chunk(chunk)
        pol.insert_audio_chunk(0, chunk)
        a = online.process_iter()
        b = pol.process_iter([0])[0]
        assert a[2] == b[2] and a[0] == b[0] and a[1] == b[1], (i, a, b)
        assert abs(online.buffer_time_offset - pol.streams[0].offset) < 1e-9
        if a[2]:
            committed.append(a)
    assert committed, "the reference policy committed nothing on the CUDA backend"
    assert online.finish()[2] == pol.finish(0)[2]
    print(f"unchanged reference policy on the GPU backend: {len(committed)} commits, "
          f"{sum(len(c[2]) for c in committed)} characters, identical to BatchedOnlineASR")


def test_unchanged_reference_dedup_and_vad_next_to_the_gpu_versions(ref):
    asr_components, _ = ref
    from audio_deduplicator import AudioDeduplicator
    from silero_vad_iterator import FakeSileroModel
    from realtime_whisper_asr_b200.dedup import B200AudioDeduplicator
    from realtime_whisper_asr_b200.vad import BatchedVADGate
    base = [speech_like(900 + i, 1.0) for i in range(5)]
    seq = [base[0], base[1], base[0], base[2], base[2] * 0.999, base[3], base[1], base[4], base[4]]
    r, g = AudioDeduplicator(), B200AudioDeduplicator()
    for k, a in enumerate(seq):
        dr = r.should_skip(a, 16000, current_time=100.0 + 0.2 * k)
        dg = g.should_skip(a, 16000, current_time=100.0 + 0.2 * k)
        assert dr[0] == dg[0] and dr[1] == dg[1], (k, dr, dg)
        if dr[2]:
            assert abs(dr[2]["similarity"] - dg[2]["similarity"]) <= 1e-4
    assert r.get_stats()["skipped_duplicate"] == g.get_stats()["skipped_duplicate"] >= 1
    x = np.concatenate([speech_like(5, 1.5) * 3, np.zeros(12000, np.float32), speech_like(6, 1.0) * 3,
                        np.zeros(16000, np.float32)]).clip(-1, 1).astype(np.float32)
    vr = asr_components.DynamicVADIterator(FakeSileroModel(0.02), initial_silence_ms=500, threshold=0.5)
    vg = BatchedVADGate(["s"], initial_silence_ms=500, threshold=0.5, energy_k=0.02)
    n_ev = 0
    for i in range(0, len(x), 640):
        a, b = vr(x[i:i + 640]), vg({"s": x[i:i + 640]})["s"]
        assert (None if a is None else {k: int(v) for k, v in a.items()}) == b, (i, a, b)
        n_ev += a is not None
    assert n_ev >= 3
