"""GPU parity of the Whisper engine (encoder, decoder logits, greedy / beam search, word timing,
language detection) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import align as OA
from oracle import decoding as OD
from oracle.model import OracleWhisper
from realtime_whisper_asr_b200 import weights as W
from realtime_whisper_asr_b200.engine import DecodeOptions, WhisperEngine
from tests.helpers import (decode_option_kwargs, mel_windows, oracle_dims, special_tokens, windows_to_device)

pytestmark = pytest.mark.gpu


class Fixture:
    def __init__(self, size, device, B, seed=0, peaked=True):
        self.md = W.MODEL_DIMS[size]
        self.sd = W.synthetic_state_dict(self.md, seed=seed, peaked=peaked)
        self.engine = WhisperEngine(self.md, W.pack_for_engine(self.sd, self.md, device), max_batch=B, max_beam=5)
        self.o32 = OracleWhisper(self.sd, oracle_dims(self.md), "fp32")
        self.o16 = OracleWhisper(self.sd, oracle_dims(self.md), "bf16")
        self.sp = special_tokens(self.md.n_vocab)
        self.mel = mel_windows([100 + i for i in range(B)], [30.0, 4.0, 11.5, 1.0, 20.0, 2.5][:B], self.md.n_mels)
        self.windows = windows_to_device(self.mel, device)
        self.enc = self.engine.encode(self.windows)
        torch.cuda.synchronize()
        melt = torch.from_numpy(self.mel)
        self.enc32 = self.o32.encode(melt)
        self.enc16 = self.o16.encode(melt)
        self.ckv = self.engine.cross_kv(self.enc)

    def prompt(self, prev=(), lang=0):
        sp = self.sp
        p = [sp["sot"], sp["lang0"] + lang, sp["transcribe"]]
        return ([sp["sot_prev"]] + list(prev) + p) if prev else p


@pytest.fixture(scope="module")
def micro(cuda_device):
    return Fixture("micro", cuda_device, B=4)


@pytest.fixture(scope="module")
def tiny(cuda_device):
    return Fixture("tiny", cuda_device, B=3)


@pytest.fixture(scope="module")
def tiny_flat(cuda_device):
    """Plain random weights: flat next-token distributions, for activation / logit tolerances."""
    return Fixture("tiny", cuda_device, B=3, peaked=False)


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm()))


@pytest.mark.parametrize("which", ["micro", "tiny", "tiny_flat"])
def test_encoder_parity(which, request):
    fx = request.getfixturevalue(which)
    got = fx.enc.float().cpu()
    # vs the bf16-contract oracle: same rounding points, only accumulation order differs
    e16 = (got - fx.enc16).abs().max().item()
    # vs pure fp32: bf16 tolerance, stated: max abs <= 0.08 (LayerNorm outputs, O(1)), cosine >= 0.9995
    e32 = (got - fx.enc32).abs().max().item()
    cos = _cos(got, fx.enc32)
    print(f"{which}: encoder max|err| vs bf16-oracle {e16:.4g}, vs fp32-oracle {e32:.4g}, cosine {cos:.6f}")
    assert e16 <= 0.05, e16
    assert e32 <= 0.08 and cos >= 0.9995, (e32, cos)


@pytest.mark.parametrize("which", ["micro", "tiny", "tiny_flat"])
def test_decoder_logits_parity(which, request):
    fx = request.getfixturevalue(which)
    rng = np.random.default_rng(5)
    seqs = [fx.prompt() + rng.integers(0, 50000, size=n).tolist() for n in (5, 17, 1, 40)][: fx.enc.shape[0]]
    got = fx.engine.decoder_logits(fx.ckv, seqs)
    torch.cuda.synchronize()
    for b, s in enumerate(seqs):
        ref16, _ = fx.o16.teacher_forced_logits(s, fx.enc16[b])
        ref32, _ = fx.o32.teacher_forced_logits(s, fx.enc32[b])
        g = got[b].cpu()
        scale = ref32.std().item()
        e16 = (g - ref16).abs().max().item() / scale
        e32 = (g - ref32).abs().max().item() / scale
        print(f"{which} seq{b}: logits err/std vs bf16-oracle {e16:.4g}, vs fp32-oracle {e32:.4g}")
        # bf16 rounding chains amplify any ulp-level difference up to the bf16 noise floor, so
        # the bf16-contract oracle is not closer than the fp32 one.  Tolerance (max over all
        # positions and the whole vocabulary, relative to the logit spread): 5 % for the plain
        # random model, 8 % for the peaked one (its last MLP multiplies the rounding of the
        # LayerNorm output by gamma = 3)
        tol = 0.05 if which.endswith("flat") else 0.08
        assert e16 <= tol, e16
        assert e32 <= tol, e32


def _decode_both(fx, beam, max_new, prev=(), margins=None):
    B = fx.enc.shape[0]
    kw = decode_option_kwargs(fx.md.n_vocab, beam_size=beam, max_new_tokens=max_new)
    # different language tokens / previous text steer the model into different transcripts
    prompts = [fx.prompt(tuple(x + 13 * b for x in prev), lang=7 * b) for b in range(B)]
    got = fx.engine.decode(fx.ckv, prompts, DecodeOptions(**kw))
    ref = []
    for b in range(B):
        m = [] if (margins is not None and beam == 1) else None
        ref.append(OD.decode_one(fx.o16, fx.enc16[b], prompts[b], OD.DecodeOptions(**kw), m))
        if margins is not None:
            margins.append(min(m) if m else None)
    return got, ref


@pytest.mark.parametrize("which,beam", [("micro", 1), ("micro", 5), ("tiny", 1), ("tiny", 5)])
def test_decode_tokens_match_oracle(which, beam, request):
    """Engine decode vs oracle decode on the same encoder input.  Greedy: identical on every chunk whose smallest
    decision margin (oracle/decoding.py) is above the bf16 noise (0.15 nats); a chunk below it is a coin flip for
    any bf16 engine and is only listed.  Beam 5 (reported, not the north-star bar): rankings of near-tied hypotheses
    may legitimately flip, so at most one chunk may differ.  The large sweep is tests/test_gpu_parity_sweep.py."""
    fx = request.getfixturevalue(which)
    margins = []
    got, ref = _decode_both(fx, beam, 0, prev=(311, 422, 533, 644) if beam == 1 else (), margins=margins)
    same, bad = 0, 0
    for b, (g, r) in enumerate(zip(got, ref)):
        ok = g.tokens == r.tokens
        same += ok
        cond = beam != 1 or margins[b] is None or margins[b] >= 0.15
        if not ok:
            n = next((i for i, (x, y) in enumerate(zip(g.tokens, r.tokens)) if x != y), min(len(g.tokens), len(r.tokens)))
            print(f"{which} beam{beam} chunk {b} (margin {margins[b]}): diverges at token {n}: {g.tokens[n:n+3]} vs {r.tokens[n:n+3]}")
            bad += cond
        assert abs(g.no_speech_prob - r.no_speech_prob) <= 1e-3 + 0.05 * r.no_speech_prob
        if ok:
            assert abs(g.sum_logprob - r.sum_logprob) <= 0.15 + 0.05 * abs(r.sum_logprob)
    print(f"{which} beam{beam}: {same}/{len(got)} chunks token-identical, lengths {[len(g.tokens) for g in got]}, margins {margins}")
    assert bad <= (0 if beam == 1 else 1), f"{bad} well-conditioned chunks differ from the oracle"


def test_timestamp_rules_hold(tiny_flat):
    tiny = tiny_flat
    got, _ = _decode_both(tiny, 1, 60)
    tb, eot = tiny.sp["timestamp_begin"], tiny.sp["eot"]
    for g in got:
        t = g.tokens
        assert t and t[0] >= tb and t[0] <= tb + 50  # first token: timestamp <= 1.0 s
        stamps = [x for x in t if x >= tb]
        assert stamps == sorted(stamps)
        assert tiny.sp["no_timestamps"] not in t and eot not in t
        for i in range(2, len(t)):
            assert not (t[i] >= tb and t[i - 1] >= tb and t[i - 2] >= tb)  # never three in a row


def test_align_matches_oracle(tiny):
    """Word-timing pass on the synthetic model's peaked alignment heads (weights._install_alignment_heads): the DTW
    path -- index work -- must EQUAL the oracle's, the token probabilities agree to bf16 tolerance."""
    fx = tiny
    sp = fx.sp
    sot_seq = [sp["sot"], sp["lang0"], sp["transcribe"]]
    scripts = W.synthetic_transcripts(fx.md.n_vocab)
    texts = [[t for t in scripts[j][:n] if t < sp["eot"]] for j, n in ((2, 40), (28, 9), (45, 70))]
    seqs = [sot_seq + [sp["no_timestamps"]] + t + [sp["eot"]] for t in texts]
    num_frames = [3000, 400, 1150]
    heads = W.ALIGNMENT_HEADS["tiny"]
    got = fx.engine.align(fx.ckv, seqs, num_frames, len(sot_seq), heads)
    for b in range(3):
        ti, fi, tp = OA.align(fx.o16, fx.enc16[b], sot_seq, sp["no_timestamps"], sp["eot"], texts[b],
                              num_frames[b], heads)
        gti, gfi, gtp = got[b]
        print(f"align seq{b}: path len {len(gti)} vs {len(ti)}, |prob err| {np.abs(gtp - tp).max():.3g}")
        assert np.allclose(gtp, tp, rtol=0.1, atol=1e-6)
        assert np.array_equal(gti, ti) and np.array_equal(gfi, fi), f"seq {b}: DTW path differs from the oracle's"
        # and the path follows the heads' ridge: one token every weights.ALIGN_STRIDE frames (+- the peak width)
        jumps = gfi[np.pad(np.diff(gti), (1, 0), constant_values=1).astype(bool)]
        inside = [j for i, j in enumerate(jumps) if 2 <= i and W.ALIGN_STRIDE * (i + 5) < num_frames[b] // 2]
        steps = np.diff(inside)
        assert len(steps) == 0 or abs(np.median(steps) - W.ALIGN_STRIDE) <= 2


def test_detect_language(tiny):
    sp = tiny.sp
    ids = list(range(sp["lang0"], sp["lang0"] + sp["n_lang"]))
    got = tiny.engine.detect_language(tiny.ckv, sp["sot"], ids)
    ref = OD.detect_language(tiny.o16, tiny.enc16, sp["sot"], ids)
    assert np.allclose(got.sum(axis=1), 1.0, atol=1e-4)
    assert np.abs(got - ref).max() <= 2e-2
    assert (got.argmax(axis=1) == ref.argmax(axis=1)).all()
