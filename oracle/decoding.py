"""Oracle (test infrastructure): greedy / beam search with Whisper's logit rules, on the CPU.

Restates CTranslate2's Whisper.generate as driven by faster-whisper
(third-party, un-vendored; call site /root/reference/asr_components.py:279-288; semantics:
SURVEY.md Appendix A.4/A.5 and OpenAI whisper/decoding.py ApplyTimestampRules / SuppressBlank /
SuppressTokens).  PARITY UNPINNED by the reference.

Search semantics (identical in the CUDA engine, csrc/engine.cu):
  * log-softmax over the admissible tokens; if the summed probability of all timestamp tokens
    exceeds the largest text-token probability, only timestamps stay admissible;
  * greedy: argmax (ties -> lower id);  beam K: per live hypothesis the top 2K tokens, merged by
    cumulative log-prob (ties -> lower hypothesis index, then lower id), the best 2K candidates are
    walked in order: EOT candidates finish (until round(K*patience) are collected), the others
    refill the beam up to K; score = cum_logprob / max(len_without_eot, 1) ** length_penalty;
  * a chunk stops when its beam is empty, enough hypotheses finished, or the length budget is
    spent (live hypotheses then finish without EOT).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np
import torch

from .model import OracleWhisper


@dataclass
class DecodeOptions:
    beam_size: int = 5
    patience: float = 1.0
    length_penalty: float = 1.0
    max_new_tokens: int = 0
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp_index: int = 50
    eot: int = 50257
    sot: int = 50258
    no_speech: int = 50362
    no_timestamps: int = 50363
    timestamp_begin: int = 50364
    suppress_ids: Sequence[int] = field(default_factory=list)
    blank_ids: Sequence[int] = field(default_factory=list)


@dataclass
class DecodeResult:
    tokens: List[int]
    sum_logprob: float
    no_speech_prob: float


def rule_mask(n_vocab: int, sampled: Sequence[int], o: DecodeOptions) -> np.ndarray:
    """Boolean mask of tokens that may NOT be sampled next (before the timestamp-mass rule)."""
    m = np.zeros(n_vocab, dtype=bool)
    ids = [i for i in o.suppress_ids if 0 <= i < n_vocab]
    m[ids] = True
    n = len(sampled)
    if n == 0 and o.suppress_blank:
        m[[i for i in o.blank_ids if 0 <= i < n_vocab]] = True
    if o.without_timestamps:
        return m
    tb = o.timestamp_begin
    m[o.no_timestamps] = True
    if n == 0:
        m[:tb] = True  # first sampled token must be a timestamp ...
        m[tb + o.max_initial_timestamp_index + 1:] = True  # ... not later than max_initial_timestamp
        return m
    last_ts = sampled[-1] >= tb
    pen_ts = n < 2 or sampled[-2] >= tb
    if last_ts:
        if pen_ts:
            m[tb:] = True  # timestamps come in pairs: text must follow
        else:
            m[:o.eot] = True  # a lone timestamp must be closed (or the chunk ended)
    stamps = [t for t in sampled if t >= tb]
    if stamps:
        lo = stamps[-1] if (last_ts and not pen_ts) else stamps[-1] + 1
        m[tb:lo] = True  # timestamps never decrease
    return m


def masked_logprobs(logits: np.ndarray, sampled: Sequence[int], o: DecodeOptions, margins: list = None) -> np.ndarray:
    """Log-probabilities after the logit rules.  ``margins`` (optional list) receives the decision margins of
    this position in nats: |log sum p(timestamps) - log max p(text)| when the timestamp-mass rule is evaluated
    with both sides admissible, and the gap between the two best admissible tokens -- what a conditioning
    analysis needs to tell a robust argmax from a near-tie."""
    x = logits.astype(np.float32).copy()
    m = rule_mask(len(x), sampled, o)
    x[m] = -np.inf
    mx = np.max(x)
    e = np.exp(x - mx, dtype=np.float32)
    lse = mx + np.log(e.sum(dtype=np.float32))
    if not o.without_timestamps:
        tb = o.timestamp_begin
        sum_ts = e[tb:].sum(dtype=np.float32)
        if sum_ts > 0:
            ts_lp = mx + np.log(sum_ts) - lse
            text_lp = np.max(x[:tb]) - lse
            if margins is not None and np.isfinite(text_lp):
                margins.append(float(abs(ts_lp - text_lp)))
            if ts_lp > text_lp:
                x[:tb] = -np.inf
                lse = mx + np.log(sum_ts)
    if margins is not None:
        top2 = np.partition(x, -2)[-2:]
        if np.isfinite(top2[0]):
            margins.append(float(top2[1] - top2[0]))
    return x - lse


def _topk(lp: np.ndarray, k: int):
    idx = np.argsort(-lp, kind="stable")[:k]
    return [(int(i), float(lp[i])) for i in idx if np.isfinite(lp[i])]


def _norm(cum: float, n: int, lp: float) -> float:
    return cum / (max(n, 1) ** lp)


def decode_one(model: OracleWhisper, enc_row: torch.Tensor, prompt: Sequence[int], o: DecodeOptions,
               margins: list = None) -> DecodeResult:
    """``margins`` (greedy only): per sampled position, the smallest decision margin (see masked_logprobs)."""
    D = model.dims
    K = o.beam_size
    k_top = 1 if K == 1 else 2 * K
    max_cand = max(1, int(round(K * o.patience)))
    cross = model.cross_kv(enc_row[None])
    st = model.new_state(1)
    hid, _ = model.decode_step(torch.tensor([list(prompt)], dtype=torch.long), st, cross, [0])
    lg = model.logits(hid)[0].numpy()
    P = len(prompt)
    sot_pos = list(prompt).index(o.sot) if o.sot in prompt else P - 1
    ns = torch.softmax(torch.from_numpy(lg[sot_pos]), dim=-1)[o.no_speech].item()
    max_new = D.n_text_ctx - P
    if o.max_new_tokens > 0:
        max_new = min(max_new, o.max_new_tokens)

    alive = []  # (tokens, cum, row)
    finished = []  # (tokens, cum, score)
    def _lp(row, sampled):
        if margins is None or K != 1:
            return masked_logprobs(row, sampled, o)
        mm: list = []
        out = masked_logprobs(row, sampled, o, mm)
        margins.append(min(mm) if mm else float("inf"))
        return out

    for tok, lp in _topk(_lp(lg[-1], []), k_top):
        if len(alive) >= K:
            break
        if tok == o.eot:
            if len(finished) < max_cand:
                finished.append(([], lp, _norm(lp, 0, o.length_penalty)))
            continue
        alive.append(([tok], lp, 0))
    done = not alive or len(finished) >= max_cand
    if not done and max_new <= 1:
        finished += [(t, c, _norm(c, len(t), o.length_penalty)) for t, c, _ in alive]
        alive, done = [], True
    step = 1
    while not done:
        st = model.reorder_state(st, [r for _, _, r in alive])
        toks = torch.tensor([[t[-1]] for t, _, _ in alive], dtype=torch.long)
        hid, _ = model.decode_step(toks, st, cross, [0] * len(alive))
        lg = model.logits(hid)[:, 0].numpy()
        cands = []
        for k, (t, c, _) in enumerate(alive):
            for tok, lp in _topk(_lp(lg[k], t), k_top):
                cands.append((c + float(np.float32(lp)), k, tok))
        cands.sort(key=lambda z: (-z[0], z[1], z[2]))
        nxt = []
        for score, k, tok in cands[:k_top]:
            if len(nxt) >= K:
                break
            t, c, _ = alive[k]
            if tok == o.eot:
                if len(finished) < max_cand:
                    finished.append((t, score, _norm(score, len(t), o.length_penalty)))
                continue
            nxt.append((t + [tok], score, k))
        alive = nxt
        if not alive or len(finished) >= max_cand:
            break
        if step + 1 >= max_new:
            finished += [(t, c, _norm(c, len(t), o.length_penalty)) for t, c, _ in alive]
            break
        step += 1
    best = None
    for f in finished:
        if best is None or f[2] > best[2]:
            best = f
    if best is None:
        return DecodeResult([], 0.0, ns)
    return DecodeResult(list(best[0]), float(np.float32(best[1])), ns)


def decode(model: OracleWhisper, enc: torch.Tensor, prompts: Sequence[Sequence[int]], o: DecodeOptions):
    return [decode_one(model, enc[b], prompts[b], o) for b in range(len(prompts))]


def detect_language(model: OracleWhisper, enc: torch.Tensor, sot: int, lang_ids: Sequence[int]) -> np.ndarray:
    out = []
    for b in range(enc.shape[0]):
        lg, _ = model.teacher_forced_logits([sot], enc[b])
        out.append(torch.softmax(lg[0, list(lang_ids)], dim=-1).numpy())
    return np.stack(out)
