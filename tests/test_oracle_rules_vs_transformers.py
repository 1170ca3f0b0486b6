"""CPU: the oracle's Whisper timestamp rules (oracle/decoding.py: rule_mask + the timestamp-mass rule) against an
independent implementation available offline, transformers' WhisperTimeStampLogitsProcessor (itself a restatement
of OpenAI's ApplyTimestampRules).  Same admissible token set and the same log-probabilities for random logits and
random sampled prefixes of every grammatical kind.  (The CUDA sample_kernel is then checked against the oracle.)"""
import types

import numpy as np
import pytest
import torch

from oracle import decoding as OD

transformers = pytest.importorskip("transformers")
from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor  # noqa: E402

V, EOT, NO_TS, TB = 600, 500, 507, 508   # small vocabulary with Whisper's layout: text < EOT < specials < timestamps


def _prefixes(rng):
    ts = lambda lo=0: int(TB + rng.integers(lo, 80))
    txt = lambda: int(rng.integers(0, EOT))
    out = [[]]
    t0 = ts()
    out += [[t0], [t0, txt()], [t0, txt(), txt()]]
    t1 = ts(t0 - TB)
    out += [[t0, txt(), t1], [t0, txt(), t1, t1], [t0, txt(), t1, t1, txt()], [t0, t0], [t0, txt(), t1, t1, txt(), ts(t1 - TB)]]
    return out


@pytest.mark.parametrize("seed", range(6))
def test_timestamp_rules_match_transformers(seed):
    rng = np.random.default_rng(seed)
    cfg = types.SimpleNamespace(no_timestamps_token_id=NO_TS, eos_token_id=EOT, bos_token_id=EOT,
                                max_initial_timestamp_index=50, _detect_timestamp_from_logprob=True)
    prompt = [EOT + 1, EOT + 2, EOT + 3]
    proc = WhisperTimeStampLogitsProcessor(cfg, begin_index=len(prompt))
    opts = OD.DecodeOptions(beam_size=1, suppress_blank=False, eot=EOT, sot=EOT + 1, no_speech=EOT + 5,
                            no_timestamps=NO_TS, timestamp_begin=TB, max_initial_timestamp_index=50,
                            suppress_ids=[], blank_ids=[])
    for sampled in _prefixes(rng):
        for peaked_ts in (False, True):
            logits = rng.normal(size=V).astype(np.float32) * 3
            if peaked_ts:
                logits[TB:] += 4.0  # make the timestamp mass win sometimes
            ref = proc(torch.tensor([prompt + sampled]), torch.from_numpy(logits)[None].clone())[0]
            ref_lp = torch.log_softmax(ref.float(), dim=-1).numpy()
            got = OD.masked_logprobs(logits, sampled, opts)
            assert np.array_equal(np.isfinite(got), np.isfinite(ref_lp)), (sampled, peaked_ts)
            fin = np.isfinite(got)
            assert fin.any()
            assert np.allclose(got[fin], ref_lp[fin], atol=2e-5), (sampled, peaked_ts)
