"""CPU: the oracle Whisper graph against transformers' Whisper with shared random weights."""
import numpy as np
import pytest
import torch

from oracle import align as OA
from oracle import decoding as OD
from oracle.model import Dims, OracleWhisper, convert_hf_state_dict


@pytest.fixture(scope="module")
def hf_pair():
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    torch.manual_seed(0)
    cfg = WhisperConfig(vocab_size=1200, num_mel_bins=80, d_model=128, encoder_layers=2, decoder_layers=2,
                        encoder_attention_heads=2, decoder_attention_heads=2, encoder_ffn_dim=512,
                        decoder_ffn_dim=512, max_source_positions=1500, max_target_positions=448,
                        activation_function="gelu", pad_token_id=0, bos_token_id=1, eos_token_id=2,
                        decoder_start_token_id=1, suppress_tokens=None, begin_suppress_tokens=None)
    hf = WhisperForConditionalGeneration(cfg).eval()
    for p in hf.parameters():
        if p.dim() == 1:
            torch.nn.init.normal_(p, 0.0, 0.1)
    sd = convert_hf_state_dict(hf.state_dict())
    dims = Dims(80, 128, 2, 2, 1200, 128, 2, 2)
    return hf, OracleWhisper(sd, dims, "fp32")


def test_encoder_and_decoder_match_transformers(hf_pair):
    hf, orc = hf_pair
    torch.manual_seed(1)
    mel = torch.randn(2, 80, 3000)
    toks = torch.randint(3, 1200, (2, 9))
    with torch.no_grad():
        ref_enc = hf.model.encoder(mel).last_hidden_state
        ref_logits = hf(input_features=mel, decoder_input_ids=toks).logits
        enc = orc.encode(mel)
        assert (enc - ref_enc).abs().max().item() <= 2e-4
        for b in range(2):
            lg, _ = orc.teacher_forced_logits(toks[b].tolist(), enc[b])
            assert (lg - ref_logits[b]).abs().max().item() <= 5e-4


def test_incremental_decoding_equals_full_pass(hf_pair):
    _, orc = hf_pair
    torch.manual_seed(2)
    enc = orc.encode(torch.randn(1, 80, 3000))
    seq = [5, 9, 100, 7, 300, 44]
    full, _ = orc.teacher_forced_logits(seq, enc[0])
    cross = orc.cross_kv(enc)
    st = orc.new_state(1)
    outs = []
    h, _ = orc.decode_step(torch.tensor([seq[:3]]), st, cross, [0])
    outs.append(orc.logits(h)[0])
    for t in seq[3:]:
        h, _ = orc.decode_step(torch.tensor([[t]]), st, cross, [0])
        outs.append(orc.logits(h)[0])
    inc = torch.cat(outs)
    assert (inc - full).abs().max().item() <= 1e-4


def _opts(**kw):
    base = dict(beam_size=1, eot=1100, sot=1101, no_speech=1102, no_timestamps=1103, timestamp_begin=1104,
                suppress_ids=[1101, 1102], blank_ids=[10, 1100], max_initial_timestamp_index=20)
    base.update(kw)
    return OD.DecodeOptions(**base)


def test_rule_mask_semantics():
    o = _opts()
    V, tb = 1200, 1104
    m = OD.rule_mask(V, [], o)
    assert m[:tb].all() and not m[tb:tb + 21].any() and m[tb + 21:].all()  # first: timestamp <= max_initial
    m = OD.rule_mask(V, [tb + 3], o)  # (implicit) pair complete -> text must follow
    assert m[tb:].all() and not m[50] and m[o.no_timestamps]
    m = OD.rule_mask(V, [tb + 3, 50], o)  # text: anything but earlier timestamps
    assert not m[50] and m[tb:tb + 4].all() and not m[tb + 4]
    m = OD.rule_mask(V, [tb + 3, 50, tb + 9], o)  # lone timestamp: only timestamp >= it, or EOT
    assert m[:o.eot].all() and not m[o.eot] and m[tb + 8] and not m[tb + 9]
    m = OD.rule_mask(V, [tb + 3, 50, tb + 9, tb + 9], o)  # pair closed: text only, later stamps
    assert m[tb:].all()
    m = OD.rule_mask(V, [], _opts(without_timestamps=True))
    assert m[10] and m[1100] and not m[50] and not m[tb + 50]


def test_greedy_and_beam_on_oracle(hf_pair):
    _, orc = hf_pair
    torch.manual_seed(3)
    enc = orc.encode(torch.randn(1, 80, 3000))
    g = OD.decode_one(orc, enc[0], [1101, 5, 6], _opts(max_new_tokens=12))
    b = OD.decode_one(orc, enc[0], [1101, 5, 6], _opts(beam_size=4, max_new_tokens=12))
    for r in (g, b):
        assert 1 <= len(r.tokens) <= 12 and r.tokens[0] >= 1104 and 0.0 <= r.no_speech_prob <= 1.0
        assert 1100 not in r.tokens
    # beam search never returns a worse length-normalised score than greedy
    assert b.sum_logprob / max(len(b.tokens), 1) >= g.sum_logprob / max(len(g.tokens), 1) - 1e-4


def test_dtw_and_median_filter():
    x = np.array([[0, 5, 5, 5], [5, 0, 0, 5], [5, 5, 5, 0]], dtype=np.float32)
    ti, fi = OA.dtw(x)
    assert list(zip(ti, fi)) == [(0, 0), (1, 1), (1, 2), (2, 3)]
    assert np.array_equal(OA.median_filter(np.array([[1., 9., 2., 8., 3.]]), 3), np.array([[9., 2., 8., 3., 8.]]))
    short = np.array([[1., 2., 3.]])
    assert np.array_equal(OA.median_filter(short, 7), short)  # too short: unfiltered


def test_align_path_is_monotone(hf_pair):
    _, orc = hf_pair
    torch.manual_seed(4)
    enc = orc.encode(torch.randn(1, 80, 3000))
    ti, fi, tp = OA.align(orc, enc[0], [1101, 5, 6], 1103, 1100, [20, 21, 22, 23], 600, [(1, 0), (1, 1)])
    assert ti[0] == 0 and fi[0] == 0 and ti[-1] == 4 and fi[-1] == 299
    assert np.all(np.diff(ti) >= 0) and np.all(np.diff(fi) >= 0) and len(tp) == 4 and np.all((tp >= 0) & (tp <= 1))


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_greedy_decode_matches_transformers_loop(hf_pair, seed):
    """Greedy decoding with Whisper's logit rules: the oracle's decode_one against a hand loop over transformers'
    model + transformers' own logits processors (suppress lists, begin-suppress, timestamp rules) + argmax."""
    from types import SimpleNamespace

    from transformers.generation.logits_process import (SuppressTokensAtBeginLogitsProcessor,
                                                        SuppressTokensLogitsProcessor,
                                                        WhisperTimeStampLogitsProcessor)
    hf, orc = hf_pair
    torch.manual_seed(seed)
    mel = torch.randn(1, 80, 3000)
    o = _opts(max_new_tokens=14)
    prompt = [1101, 5, 6]
    with torch.no_grad():
        enc = orc.encode(mel)
        got = OD.decode_one(orc, enc[0], prompt, o).tokens
        cfg = SimpleNamespace(no_timestamps_token_id=o.no_timestamps, eos_token_id=o.eot, bos_token_id=o.eot,
                              max_initial_timestamp_index=o.max_initial_timestamp_index,
                              _detect_timestamp_from_logprob=True)
        procs = [SuppressTokensLogitsProcessor(list(o.suppress_ids)),
                 SuppressTokensAtBeginLogitsProcessor(list(o.blank_ids), begin_index=len(prompt)),
                 WhisperTimeStampLogitsProcessor(cfg, begin_index=len(prompt))]
        enc_hf = hf.model.encoder(mel).last_hidden_state
        ids = list(prompt)
        ref = []
        for _ in range(14):
            logits = hf(encoder_outputs=(enc_hf,), decoder_input_ids=torch.tensor([ids])).logits[:, -1, :]
            for p in procs:
                logits = p(torch.tensor([ids]), logits)
            t = int(torch.argmax(logits, dim=-1))
            if t == o.eot:
                break
            ids.append(t)
            ref.append(t)
    assert got == ref


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_dtw_and_median_match_transformers(seed):
    """Word-timing building blocks against transformers' restatement of OpenAI's timing.py: same DTW path (tie
    breaking included) and the same reflect-padded median filter on random matrices."""
    g = pytest.importorskip("transformers.models.whisper.generation_whisper")
    rng = np.random.default_rng(seed)
    m = rng.normal(size=(17 + seed, 140)).astype(np.float32)
    m[3] = m[4]  # exact ties along a row pair
    ti, fi = OA.dtw(-m)
    rti, rfi = g._dynamic_time_warping(-m)
    assert np.array_equal(ti, rti) and np.array_equal(fi, rfi)
    w = rng.normal(size=(3, 9, 75)).astype(np.float32)
    assert np.array_equal(OA.median_filter(w, 7), g._median_filter(torch.from_numpy(w), 7).numpy())
