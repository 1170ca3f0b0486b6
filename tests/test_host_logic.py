"""CPU: host-side logic of the product (tokenizer, prompt, segment split, word merge) and the
no-fallback rule."""
import numpy as np
import pytest

from realtime_whisper_asr_b200 import transcribe as T
from realtime_whisper_asr_b200 import weights as W
from realtime_whisper_asr_b200.tokenizer import Tokenizer


def test_special_token_layout():
    tk = Tokenizer(51865, True, language="en")
    assert (tk.eot, tk.sot, tk.translate, tk.transcribe, tk.sot_prev, tk.no_speech, tk.no_timestamps,
            tk.timestamp_begin) == (50257, 50258, 50358, 50359, 50361, 50362, 50363, 50364)
    assert tk.sot_sequence == [50258, 50259, 50359] and len(tk.languages) == 99
    v3 = Tokenizer(51866, True, task="translate", language="yue")
    assert (v3.translate, v3.transcribe, v3.no_timestamps, v3.timestamp_begin) == (50359, 50360, 50364, 50365)
    assert v3.language == 50259 + 99 and v3.sot_sequence[-1] == 50359 and len(v3.languages) == 100
    with pytest.raises(ValueError):
        Tokenizer(51865, True, language="xx")
    assert W.special_token_ids(51866)["timestamp_begin"] == 50365


def test_byte_vocab_roundtrip_and_words():
    tk = Tokenizer(51865, True, language="en")
    text = " héllo wörld, 你好"
    assert tk.decode(tk.encode(text)) == text
    words, wt = tk.split_to_word_tokens(tk.encode(" hello big world.") + [tk.eot])
    assert [w for w in words[:-1]] == [" hello", " big", " world", "."] and wt[-1] == [tk.eot]
    zh = Tokenizer(51865, True, language="zh")
    words, wt = zh.split_to_word_tokens(zh.encode("你好") + [zh.eot])
    assert words[:2] == ["你", "好"] and [len(t) for t in wt[:2]] == [3, 3]
    assert tk.decode_with_timestamps([tk.timestamp_begin + 50, 104, 105, tk.timestamp_begin + 100]) == "<|1.00|>hi<|2.00|>"
    sup = tk.suppressed_tokens([-1])
    assert tk.sot in sup and tk.no_speech in sup and tk.eot not in sup and sup == sorted(set(sup))
    assert tk.blank_tokens() == [32, tk.eot]


def test_get_prompt():
    tk = Tokenizer(51865, True, language="en")
    assert T.get_prompt(tk, []) == tk.sot_sequence
    prev = list(range(1000, 1400))
    p = T.get_prompt(tk, prev)
    assert p[0] == tk.sot_prev and p[1:224] == prev[-223:] and p[224:] == tk.sot_sequence and len(p) == 227
    p = T.get_prompt(tk, [], without_timestamps=True, prefix="ab")
    assert p == tk.sot_sequence + [tk.no_timestamps] + tk.encode(" ab")


def test_split_segments_by_timestamps():
    tk = Tokenizer(51865, True, language="en")
    tb = tk.timestamp_begin
    toks = [tb + 0, 500, 501, tb + 100, tb + 100, 502, tb + 250]  # two closed segments, single ts ending
    segs, seek, single = T.split_segments_by_timestamps(tk, toks, 10.0, 3000, 30.0, 1000)
    assert single and seek == 4000
    assert [(s["start"], s["end"]) for s in segs] == [(10.0, 12.0), (12.0, 15.0)]
    assert segs[0]["tokens"] == toks[:4] and segs[1]["tokens"] == toks[4:]
    toks = [tb + 0, 500, tb + 100, tb + 100, 502, 503]  # text after the last pair: resume at 2.00 s
    segs, seek, single = T.split_segments_by_timestamps(tk, toks, 0.0, 3000, 30.0, 0)
    assert not single and seek == 200 and len(segs) == 1 and segs[0]["end"] == 2.0
    toks = [tb + 5, 500, 501]  # no consecutive timestamps: one segment of the whole window
    segs, seek, single = T.split_segments_by_timestamps(tk, toks, 3.0, 1200, 12.0, 300)
    assert seek == 1500 and segs[0]["start"] == 3.0 and abs(segs[0]["end"] - 3.1) < 1e-9
    segs, seek, _ = T.split_segments_by_timestamps(tk, [500, 501], 3.0, 1200, 12.0, 300)
    assert segs[0]["end"] == 15.0


def test_merge_punctuations_and_words_from_alignment():
    al = [dict(word=" (", tokens=[1]), dict(word="hello", tokens=[2]), dict(word=",", tokens=[3]),
          dict(word=" world", tokens=[4]), dict(word="!", tokens=[5])]
    T.merge_punctuations(al, "\"'“¿([{-", "\"'.。,，!！?？:：”)]}、")
    assert [a["word"] for a in al] == ["", " (hello,", "", " world!", ""]
    assert [a["tokens"] for a in al] == [[], [1, 2, 3], [], [4, 5], []]
    tk = Tokenizer(51865, True, language="en")
    text = tk.encode(" ab cd")
    # path: token i -> frames; two words of 3 byte tokens each + eot
    ti = np.array([0, 0, 1, 2, 3, 3, 4, 5, 6])
    fi = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8])
    tp = np.linspace(0.1, 0.6, 6).astype(np.float32)
    ws = T.words_from_alignment(tk, text, ti, fi, tp)
    assert [w["word"] for w in ws] == [" ab", " cd"]
    assert ws[0]["start"] == 0.0 and ws[0]["end"] == ws[1]["start"] == 4 / 50 and ws[1]["end"] == 8 / 50
    assert abs(ws[0]["probability"] - tp[:3].mean()) < 1e-6


def test_synthetic_transcripts_are_consistent():
    tr = W.synthetic_transcripts(51865)
    tb = W.special_token_ids(51865)["timestamp_begin"]
    assert len(tr) == 51
    seen = set()
    for j, seq in enumerate(tr):
        assert seq[0] == tb + j and seq[-1] >= tb
        stamps = [t for t in seq if t >= tb]
        assert stamps == sorted(stamps) and len(set(seq)) == len(seq)
        assert not seen & set(seq)
        seen |= set(seq)


def test_no_cpu_fallback():
    import torch
    from realtime_whisper_asr_b200 import _lib
    from realtime_whisper_asr_b200.backend import B200WhisperASR
    from realtime_whisper_asr_b200.transcribe import B200WhisperModel
    with pytest.raises(_lib.B200WError):
        B200WhisperModel("tiny", device="cpu")
    with pytest.raises(RuntimeError):
        B200WhisperASR("en", "tiny", device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(_lib.B200WError):
            B200WhisperModel("tiny", device="cuda")


def test_temperature_fallback_ladder_host_logic():
    """transcribe._decode_with_fallback against faster-whisper's generate_with_fallback rules, with a scripted engine
    (no GPU): which chunks are decoded again at which temperature, which result each chunk keeps."""
    import types

    import torch

    from realtime_whisper_asr_b200.engine import DecodeResult
    from realtime_whisper_asr_b200.transcribe import B200WhisperModel, TranscriptionOptions

    tk = Tokenizer(51865, True, language="en")
    good = tk.encode(" the quick brown fox jumps over the lazy dog")
    rep = tk.encode(" la" * 60)  # compression ratio far above 2.4

    def res(tokens, avg_logprob, no_speech=0.0):
        return DecodeResult(list(tokens), avg_logprob * (len(tokens) + 1), no_speech)

    # script[temperature][chunk] -> result
    script = {
        0.0: [res(good, -0.2), res(rep, -0.3), res(good, -1.5), res(good, -2.0, no_speech=0.9), res(rep, -0.9)],
        0.4: [None, res(good, -0.5), res(good, -1.4), None, res(rep, -0.5)],
        0.8: [None, None, res(good, -1.2), None, res(rep, -0.7)],
    }
    calls = []

    class Engine:
        def decode(self, ckv, prompts, opts):
            idx = [int(p[0]) for p in prompts]  # the prompt carries the chunk's index
            calls.append((round(opts.temperature, 3), idx, opts.best_of))
            assert ckv.shape[0] == len(prompts)
            return [script[round(opts.temperature, 3)][i] for i in idx]

    fake = types.SimpleNamespace(_sample_seed=0, max_beam=5, tokenizer=lambda lang, task="transcribe": tk,
                                 suppress_ids=None, suppress_ids_begin=None)
    fake._decode_options = types.MethodType(B200WhisperModel._decode_options, fake)
    o = TranscriptionOptions(5, 3, 1.0, 1.0, 1.0, 0, -1.0, 0.6, 2.4, True, 0.5, [0.0, 0.4, 0.8], None, None, True, [-1],
                             False, 1.0, False, "", "", None)
    batch = [types.SimpleNamespace(prompt=[i], tokenizer=tk) for i in range(5)]
    ckv = torch.zeros(5, 1)
    final, temps = B200WhisperModel._decode_with_fallback(fake, Engine(), ckv, batch, o)
    # rung 0: everything; rung 0.4: the repetitive (1, 4) and the unlikely (2) chunks -- not the silent one (3);
    # rung 0.8: what still fails
    assert calls == [(0.0, [0, 1, 2, 3, 4], 3), (0.4, [1, 2, 4], 3), (0.8, [2, 4], 3)]
    assert temps[0] == 0.0 and final[0] is script[0.0][0]
    assert temps[1] == 0.4 and final[1] is script[0.4][1]           # fixed by the first fallback
    assert temps[3] == 0.0 and final[3] is script[0.0][3]           # silence: low log-probability is not a failure
    assert temps[2] == 0.8 and final[2] is script[0.8][2]           # failed everywhere: best average log-probability
    assert temps[4] == 0.4 and final[4] is script[0.4][4]           # all above the compression threshold: best of all
    assert fake._sample_seed == 3                                    # one seed per decode call


def test_words_from_alignment_matches_the_numpy_statement():
    """transcribe.words_from_alignment (faster-whisper find_alignment after the DTW): the word probabilities are the
    means of their tokens' probabilities (plain Python sums since round 2), times come from the path's jumps."""
    rng = np.random.default_rng(3)
    tk = Tokenizer(51865, True, language="en")
    text = tk.encode(" hello world, this is a test of word timing.")
    n = len(text) + 1  # + EOT, as the aligner sees it
    frames = np.sort(rng.integers(0, 400, size=n)).astype(np.int64)
    text_indices = np.repeat(np.arange(n), 3)
    time_indices = np.repeat(frames, 3) + np.tile(np.arange(3), n)
    probs = rng.random(n - 1).astype(np.float32)
    words = T.words_from_alignment(tk, text, text_indices, time_indices, probs)
    w_text, w_tokens = tk.split_to_word_tokens(text + [tk.eot])
    assert [w["word"] for w in words] == w_text[:-1] and [w["tokens"] for w in words] == w_tokens[:-1]
    bounds = np.pad(np.cumsum([len(t) for t in w_tokens[:-1]]), (1, 0))
    jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
    jump_times = time_indices[jumps] / 50.0
    for k, w in enumerate(words):
        assert w["start"] == pytest.approx(jump_times[bounds[k]]) and w["end"] == pytest.approx(jump_times[bounds[k + 1]])
        assert w["probability"] == pytest.approx(float(np.mean(probs[bounds[k]:bounds[k + 1]])), rel=1e-6)
    # the cached single-token decodes give the same split as a fresh tokenizer
    assert Tokenizer(51865, True, language="en").split_to_word_tokens(text + [tk.eot]) == (w_text, w_tokens)
