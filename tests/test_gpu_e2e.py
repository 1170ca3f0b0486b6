"""GPU end-to-end: the backend class behind the streaming policy (harness stand-ins of the
un-vendored upstream processors), batching invariance, the dedup drop-in, multi-window audio."""
import os
import threading

import numpy as np
import pytest
import torch

from realtime_whisper_asr_b200 import weights as W
from realtime_whisper_asr_b200.backend import B200WhisperASR
from realtime_whisper_asr_b200.dedup import B200AudioDeduplicator, should_skip_batch
from realtime_whisper_asr_b200.scheduler import BatchedASR, StreamBatcher
from realtime_whisper_asr_b200.tokenizer import Tokenizer
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests import harness
from tests.synth import speech_like

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def model(cuda_device):
    return B200WhisperModel("tiny", max_batch=8, max_beam=5, synthetic=True)


def _tokens(segs):
    return [t for s in segs for t in s.tokens]


def test_backend_contract(model):
    asr = B200WhisperASR("en", model=model, transcribe_kwargs={"beam_size": 1, "temperature": 0.0})
    audio = speech_like(3, 8.0)
    res = asr.transcribe(audio, init_prompt="previous words")
    words = asr.ts_words(res)
    ends = asr.segments_end_ts(res)
    assert res and words and len(ends) == len(res) and asr.sep == ""
    assert all(isinstance(w[2], str) and w[1] >= w[0] >= 0.0 for w in words)
    assert [w[0] for w in words] == sorted(w[0] for w in words)            # monotone word starts
    assert ends == sorted(ends)  # (synthetic weights ignore the clip length: no bound on the last end)
    assert asr.transcribe(np.zeros(0, np.float32)) is not None               # empty audio does not raise
    asr.set_translate_task()
    assert asr.transcribe_kargs["task"] == "translate" and asr.transcribe(audio)  # translate task token path
    with pytest.raises(NotImplementedError):
        asr.use_vad()
        asr.transcribe(audio)


def test_temperature_sampling_and_fallback(model):
    """faster-whisper's temperature semantics (the reference forwards `temperature`, asr_components.py:283, and its
    adaptive hook may raise it, speech_rate_audio_processor.py:214-218): a positive temperature samples best_of
    hypotheses, a tuple is a fallback ladder driven by the compression-ratio / log-probability thresholds."""
    audio = speech_like(5, 10.0)
    kw = dict(language="en", beam_size=5, word_timestamps=True, condition_on_previous_text=False)
    greedy, _ = model.transcribe(audio, temperature=0.0, **kw)
    greedy = list(greedy)
    assert greedy and all(s.temperature == 0.0 for s in greedy)
    # sampling: well-formed (timestamp grammar, ordered words), flagged with its temperature
    a, _ = model.transcribe(audio, temperature=0.8, best_of=3, **kw)
    a = list(a)
    assert a and all(s.temperature == 0.8 for s in a)
    for s in a:
        assert s.end >= s.start and all(w.end >= w.start for w in s.words)
    # a ladder whose first rung cannot pass (log-probability threshold above 0) falls through to the last rung; every
    # rung fails, so each chunk keeps the rung with the best average log-probability
    b, info = model.transcribe(audio, temperature=(0.0, 0.5), best_of=2, log_prob_threshold=1.0,
                               no_speech_threshold=None, **kw)
    b = list(b)
    assert b and all(s.temperature in (0.0, 0.5) for s in b)
    assert info.transcription_options.temperatures == [0.0, 0.5]
    # thresholds that always pass: the ladder stops at the first rung and equals plain temperature-0 decoding
    c, _ = model.transcribe(audio, temperature=(0.0, 0.5, 1.0), log_prob_threshold=None,
                            compression_ratio_threshold=None, **kw)
    assert _tokens(list(c)) == _tokens(greedy)
    # the ASR backend forwards the keyword unchanged
    asr = B200WhisperASR("en", model=model, transcribe_kwargs={"beam_size": 5, "temperature": 0.1, "best_of": 2})
    assert asr.transcribe(audio, init_prompt="")


def test_model_dir_formats_and_missing_checkpoint(cuda_device, tmp_path):
    """the reference constructs its backend from a CTranslate2 model directory or a size name
    (/root/reference/asr_components.py:232-265).  A CT2 model.bin directory and a Hugging Face safetensors directory
    of the same weights give the engine bit-identical parameters (same encoder output, same tokens) as the state
    dict they were written from; a size name with no checkpoint on disk is an error, never silent random weights."""
    from tests import ct2_format as C
    md = W.MODEL_DIMS["tiny"]
    sd = W.synthetic_state_dict(md, seed=0)
    heads = W.alignment_heads_for(md)
    from realtime_whisper_asr_b200 import frontend as FE
    from realtime_whisper_asr_b200.engine import DecodeOptions
    from tests.helpers import decode_option_kwargs
    ref = B200WhisperModel("tiny", state_dict=sd, dims=md, max_batch=2, max_beam=1)
    audio = speech_like(11, 6.0)
    dev, offs = FE.upload_audio([audio], cuda_device)
    win = FE.pack_windows(FE.log_mel(dev, offs, md.n_mels), [0], [0])
    opts = DecodeOptions(**decode_option_kwargs(md.n_vocab, beam_size=1))
    prompt = [ref.tokenizer("en").sot_sequence]
    enc_ref = ref.engine.encode(win)
    tok_ref = ref.engine.decode(ref.engine.cross_kv(enc_ref), prompt, opts)[0].tokens
    assert len(tok_ref) > 20
    root = tmp_path / "models_fast"
    d1 = C.write_ct2_dir(str(root / "models--Systran--faster-whisper-tiny" / "snapshots" / "r0"), sd, md,
                         dtype="float32", alignment_heads=heads, suppress_ids=[1, 2, 7], suppress_ids_begin=[220, 50257])
    d2 = C.write_hf_safetensors_dir(str(tmp_path / "hf"), sd, alignment_heads=heads)
    for kwargs in (dict(model_size_or_path="tiny", download_root=str(root)), dict(model_size_or_path=d2)):
        m = B200WhisperModel(max_batch=2, max_beam=1, **kwargs)
        assert not m.synthetic and m.alignment_heads == heads and m.tokenizer_json.endswith("tokenizer.json")
        enc = m.engine.encode(win)
        assert torch.equal(enc, enc_ref)                       # identical parameters -> identical arithmetic
        assert m.engine.decode(m.engine.cross_kv(enc), prompt, opts)[0].tokens == tok_ref
        segs, info = m.transcribe(audio, language="en", beam_size=1, word_timestamps=True)  # HF vocabulary path runs
        assert info.language == "en" and isinstance(list(segs), list)
        if "download_root" in kwargs:  # CT2 config.json lists replace the vocabulary-derived suppress sets
            o = m._decode_options(m.tokenizer("en"), m.transcribe_batch([audio], languages=["en"], beam_size=1, temperature=0.0)[0][1].transcription_options)
            assert {1, 2, 7} <= set(o.suppress_ids) and list(o.blank_ids) == [220, 50257]
        m.engine.close()
    # through the reference-shaped constructor too: modelsize + cache_dir (MAIN:1125-1134)
    asr = B200WhisperASR("en", "tiny", cache_dir=str(root), transcribe_kwargs={"beam_size": 1}, max_batch=1)
    # (the test tokenizer.json has ~360 pieces: most synthetic ids decode to nothing, so segments may be empty)
    assert isinstance(asr.transcribe(audio), list) and not asr.model.synthetic
    with pytest.raises(FileNotFoundError, match="no checkpoint"):
        B200WhisperModel("base", download_root=str(root))
    with pytest.raises(FileNotFoundError, match="no checkpoint"):
        B200WhisperASR("en", "base", cache_dir=str(root))
    assert d1


def test_options_degrade_instead_of_raising(model):
    """/root/reference/asr_components.py:270-288 forwards whatever beam_size / temperature its config or adaptive
    hook holds (speech_rate_audio_processor.py:214-218: beam up to 10, temperature 0.1).  Neither may throw in the
    middle of a stream: the beam (and best_of) is clamped to the engine's, a positive temperature samples;
    token-id initial prompts (faster-whisper's Iterable[int] form) are accepted."""
    audio = speech_like(13, 5.0)
    base = [s.tokens for s in model.transcribe(audio, language="en", beam_size=5, temperature=0.0)[0]]
    assert [s.tokens for s in model.transcribe(audio, language="en", beam_size=10, temperature=0.0)[0]] == base
    assert list(model.transcribe(audio, language="en", beam_size=10, best_of=10, temperature=0.1)[0])
    tk = model.tokenizer("en")
    a = [s.tokens for s in model.transcribe(audio, language="en", beam_size=1, temperature=0.0,
                                            initial_prompt="hello there")[0]]
    b = [s.tokens for s in model.transcribe(audio, language="en", beam_size=1, temperature=0.0,
                                            initial_prompt=tk.encode(" hello there"))[0]]
    assert a == b
    two = model.transcribe_batch([audio, audio], languages=["en", "en"], beam_size=1, temperature=0.0,
                                 initial_prompts=["hello there", None])
    assert [s.tokens for s in two[0][0]] == a


def test_batch_invariance_and_multi_window(model):
    auds = [speech_like(40 + i, s) for i, s in enumerate([30.0, 47.0, 3.0, 12.0, 65.0])]
    kw = dict(beam_size=5, temperature=0.0, word_timestamps=True, condition_on_previous_text=True)
    batch = model.transcribe_batch(auds, languages=["en"] * 5, **kw)
    for a, (segs_b, info_b) in zip(auds, batch):
        segs_s, info_s = model.transcribe(a, language="en", **kw)
        segs_s = list(segs_s)
        assert _tokens(segs_s) == _tokens(segs_b)
        assert [(s.start, s.end) for s in segs_s] == [(s.start, s.end) for s in segs_b]
        assert abs(info_b.duration - len(a) / 16000) < 1e-6
    long_segs = batch[4][0]
    assert long_segs[-1].end > 30.0 and max(s.seek for s in long_segs) >= 2000  # several windows were decoded
    # (no ordering claim across windows: the synthetic model ignores the audio, so a window's
    # first word may start before the previous window's last segment ended)


def test_concurrent_lanes_give_identical_results(model):
    """lanes > 1: the batch runs as concurrent sub-batches on separate engines / CUDA streams / host threads.
    Streams are independent and every kernel is batch-invariant, so tokens, word times and probabilities are
    identical to the single-lane result -- also when the lanes are unevenly filled or streams span several windows."""
    auds = [speech_like(140 + i, s) for i, s in enumerate([30.0, 47.0, 3.0, 12.0, 65.0, 1.0, 22.0])]
    kw = dict(beam_size=5, temperature=0.0, word_timestamps=True, condition_on_previous_text=True)
    want = model.transcribe_batch(auds, languages=["en"] * len(auds), **kw)
    sd = W.synthetic_state_dict(model.dims, seed=0)
    for lanes in (2, 3):
        m = B200WhisperModel("tiny", state_dict=sd, dims=model.dims, max_batch=8, max_beam=5, lanes=lanes)
        assert m.n_lanes == lanes and len(m.engines) == lanes
        for _ in range(2):
            got = m.transcribe_batch(auds, languages=["en"] * len(auds), **kw)
            for (sa, _), (sb, _) in zip(want, got):
                assert [s.tokens for s in sa] == [s.tokens for s in sb]
                assert [(w.start, w.end, w.word, w.probability) for s in sa for w in s.words] == \
                       [(w.start, w.end, w.word, w.probability) for s in sb for w in s.words]
        for e in m.engines:
            e.close()


def test_multi_gpu_model_one_process(model):
    """MultiGpuModel: one process, one replica per GPU, streams sharded by key.  Needs >= 2 GPUs (gpurun --gpus 2);
    results are identical to the single-GPU model, and both devices actually ran kernels (per-device kernel
    attributes and SM counts: a second GPU in one process used to fail to launch)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from realtime_whisper_asr_b200.scheduler import MultiGpuModel
    from realtime_whisper_asr_b200.streaming import BatchedOnlineASR
    sd = W.synthetic_state_dict(model.dims, seed=0)
    mg = MultiGpuModel(lambda d: B200WhisperModel("tiny", state_dict=sd, dims=model.dims, device_index=d, max_batch=4,
                                                  max_beam=5), [0, 1])
    auds = [speech_like(240 + i, s) for i, s in enumerate([30.0, 7.0, 3.0, 12.0, 45.0, 1.0, 22.0])]
    kw = dict(beam_size=5, temperature=0.0, word_timestamps=True, condition_on_previous_text=True)
    want = model.transcribe_batch(auds, languages=["en"] * len(auds), **kw)
    got = mg.transcribe_batch(auds, languages=["en"] * len(auds), **kw)
    for (sa, _), (sb, _) in zip(want, got):
        assert [s.tokens for s in sa] == [s.tokens for s in sb]
        assert [(w.start, w.end) for s in sa for w in s.words] == [(w.start, w.end) for s in sb for w in s.words]
    assert all(c > 0 for c in mg.calls_per_device)
    # the batched streaming policy runs on top of it unchanged
    pol = BatchedOnlineASR(mg, transcribe_kwargs={"beam_size": 1})
    for i in range(6):
        pol.add_stream(i, "en")
    for it in range(3):
        for i in range(6):
            pol.insert_audio_chunk(i, auds[0][(it * 6 + i) * 8000:(it * 6 + i + 2) * 8000])
        out = pol.process_iter()
        assert len(out) == 6
    mg.close()


def test_language_detection_and_info(model):
    segs, info = model.transcribe(speech_like(5, 4.0), language=None, beam_size=1, temperature=0.0)
    assert info.language in model.supported_languages and 0.0 < info.language_probability <= 1.0
    assert len(info.all_language_probs) == 99 and abs(sum(p for _, p in info.all_language_probs) - 1.0) < 1e-3
    lang, prob, allp = model.detect_language(speech_like(5, 4.0))
    assert lang == info.language


def test_online_processor_fixed_interval(model):
    """BASELINE config 1 shape: one 30 s stream in 1 s increments through OnlineASRProcessor, greedy."""
    harness.install()
    from whisper_online import OnlineASRProcessor
    asr = B200WhisperASR("en", model=model, transcribe_kwargs={"beam_size": 1, "temperature": 0.0})
    online = OnlineASRProcessor(asr, buffer_trimming=("segment", 15))
    audio = speech_like(77, 30.0)
    out = []
    for i in range(30):
        online.insert_audio_chunk(audio[i * 16000:(i + 1) * 16000])
        beg, end, text = online.process_iter()
        if text:
            out.append((beg, end, text))
            assert end >= beg >= 0
    tail = online.finish()
    assert out, "nothing was committed in 30 s"
    begs = [b for b, _, _ in out]
    assert begs == sorted(begs)
    assert len(online.audio_buffer) / 16000 <= 30.0


def test_vac_processor_with_fake_vad(model, monkeypatch):
    harness.install()
    import whisper_online
    from silero_vad_iterator import FakeSileroModel
    monkeypatch.setattr(torch.hub, "load", lambda *a, **k: (FakeSileroModel(), None))
    asr = B200WhisperASR("en", model=model, transcribe_kwargs={"beam_size": 1})
    vac = whisper_online.VACOnlineASRProcessor(0.5, asr, buffer_trimming=("segment", 15))
    audio = np.concatenate([speech_like(8, 4.0), np.zeros(16000 * 2, np.float32), speech_like(9, 3.0)])
    texts = []
    for i in range(0, len(audio), 640):  # 0.04 s reads, like the reference loop
        vac.insert_audio_chunk(audio[i:i + 640])
        if (i // 640) % 12 == 11:
            b, e, t = vac.process_iter()
            if t:
                texts.append(t)
    assert texts, "VAC path produced no text"


def test_stream_batcher_matches_solo_calls(model):
    b = StreamBatcher(model, max_batch=8, max_wait_ms=30)
    auds = [speech_like(60 + i, 4.0 + i) for i in range(6)]
    got = {}

    def run(i):
        asr = BatchedASR("en", b, transcribe_kwargs={"beam_size": 1})
        got[i] = asr.transcribe(auds[i], init_prompt="abc" if i % 2 else "")

    ts = [threading.Thread(target=run, args=(i,)) for i in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    b.close()
    solo = B200WhisperASR("en", model=model, transcribe_kwargs={"beam_size": 1})
    for i in range(6):
        ref = solo.transcribe(auds[i], init_prompt="abc" if i % 2 else "")
        assert _tokens(got[i]) == _tokens(ref)
    assert b.batches_run < 6  # concurrent calls were actually batched


def test_dedup_dropin_matches_reference_golden(cuda_device):
    """Decisions of should_skip on a stream with repeats: golden from the reference class."""
    g = np.load(os.path.join(GOLD, "dedup_golden.npz"))
    d = B200AudioDeduplicator()
    base = [speech_like(900 + i, 1.0) for i in range(4)]
    seq = [base[0], base[1], base[0], base[2], base[2] * 0.999, base[3], base[1]]
    decisions, sims = [], []
    for k, a in enumerate(seq):
        skip, reason, details = d.should_skip(a, 16000, current_time=100.0 + 0.2 * k)
        decisions.append(int(skip))
        sims.append(details["similarity"] if details else -1.0)
    assert decisions == g["decisions"].tolist()
    for a, b in zip(sims, g["sims"].tolist()):
        assert abs(a - b) <= 1e-4
    assert d.get_stats()["total_checked"] == 7
    # batched variant: one launch for many streams
    ds = [B200AudioDeduplicator() for _ in range(3)]
    r1 = should_skip_batch(ds, [base[0], base[1], np.zeros(100, np.float32)], current_time=1.0)
    r2 = should_skip_batch(ds, [base[0], base[2], base[3]], current_time=1.5)
    assert [x[0] for x in r1] == [False, False, False] and r2[0][0] is True and r2[0][1] == "duplicate"
