"""Generate the token / word-timing parity fixtures: the CPU oracle's greedy transcripts and DTW paths of seeded
synthetic chunks, per model size, written to tests/golden/parity_<size>.json (+ .npz for the paths).

    python tools/make_parity_golden.py tiny 48        # size, number of chunks
    python tools/make_parity_golden.py large-v3 16

The chunk list is a pure function of the index k (tests/parity_spec.py): audio seed, duration (0.5 .. 30 s),
language and init_prompt text (0 / ~50 / ~300 characters), i.e. the call shape of
/root/reference/asr_components.py:279-288.  Per chunk the oracle (oracle/, "bf16" precision contract = what a bf16
engine materialises between ops) computes log-mel -> first 30 s window -> encoder -> greedy decode with the Whisper
logit rules, and records
  tokens       the decoded ids of the window,
  margin       the smallest decision margin over all sampled positions (nats; oracle/decoding.py masked_logprobs):
               a chunk whose margin is below the engine-vs-oracle noise is a coin flip for ANY bf16 engine, so the
               GPU test reports such chunks separately (they are listed, not hidden),
  path         the DTW path (text index, frame index) of the word-timing pass over the decoded text tokens.
Weights: realtime_whisper_asr_b200.weights.synthetic_state_dict(dims, seed=0) generated on the CPU (deterministic),
the same call the GPU test makes.  Nothing here touches the CUDA path.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import align as OA  # noqa: E402
from oracle import decoding as OD  # noqa: E402
from oracle import frontend as OF  # noqa: E402
from oracle.model import OracleWhisper  # noqa: E402
from realtime_whisper_asr_b200 import weights as W  # noqa: E402
from realtime_whisper_asr_b200.tokenizer import Tokenizer  # noqa: E402
from tests.helpers import oracle_dims  # noqa: E402
from tests.parity_spec import chunk_spec, prompt_tokens  # noqa: E402
from tests.synth import speech_like  # noqa: E402


def main():
    size, n = sys.argv[1], int(sys.argv[2])
    beam = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    out_json = os.path.join(ROOT, "tests", "golden", f"parity_{size}{'' if beam == 1 else f'_beam{beam}'}.json")
    out_npz = out_json.replace(".json", ".npz")
    md = W.MODEL_DIMS[size]
    t0 = time.time()
    sd = W.synthetic_state_dict(md, seed=0)
    orc = OracleWhisper(sd, oracle_dims(md), "bf16")
    heads = W.alignment_heads_for(md)
    print(f"{size}: weights in {time.time() - t0:.1f} s", flush=True)
    chunks, paths = [], {}
    for k in range(n):
        sp = chunk_spec(k)
        t0 = time.time()
        tk = Tokenizer(md.n_vocab, True, language=sp["language"])
        audio = speech_like(sp["seed"], sp["seconds"])
        feats = OF.log_mel(audio, md.n_mels)
        content_frames = feats.shape[-1] - 1
        seg = min(3000, content_frames)
        mel = OF.pad_or_trim(feats[:, :3000])
        enc = orc.encode(torch.from_numpy(mel[None]))
        prompt = prompt_tokens(tk, sp["prompt"])
        opts = OD.DecodeOptions(beam_size=beam, eot=tk.eot, sot=tk.sot, no_speech=tk.no_speech,
                                no_timestamps=tk.no_timestamps, timestamp_begin=tk.timestamp_begin,
                                suppress_ids=tk.suppressed_tokens([-1]), blank_ids=tk.blank_tokens())
        margins = []
        res = OD.decode_one(orc, enc[0], prompt, opts, margins if beam == 1 else None)
        text = [t for t in res.tokens if t < tk.eot]
        rec = dict(k=k, **sp, tokens=res.tokens, sum_logprob=res.sum_logprob, no_speech_prob=res.no_speech_prob,
                   margin=float(min(margins)) if margins else None,
                   first_margin=float(margins[0]) if margins else None, segment_size=int(seg))
        if text and beam == 1:
            ti, fi, tp = OA.align(orc, enc[0], tk.sot_sequence, tk.no_timestamps, tk.eot, text, seg, heads)
            paths[f"ti{k}"] = ti.astype(np.int16)
            paths[f"fi{k}"] = fi.astype(np.int16)
            paths[f"tp{k}"] = tp.astype(np.float32)
        chunks.append(rec)
        print(f"  chunk {k}: {sp['seconds']:5.1f} s {sp['language']} prompt {len(sp['prompt']):3d} chars -> "
              f"{len(res.tokens)} tokens, first ts {res.tokens[0] - tk.timestamp_begin if res.tokens else None}, "
              f"margin {rec['margin']}, {time.time() - t0:.1f} s", flush=True)
        with open(out_json, "w") as f:
            json.dump(dict(size=size, beam=beam, weights="synthetic_state_dict(seed=0, cpu)", precision="bf16",
                           chunks=chunks), f)
        if paths:
            np.savez_compressed(out_npz, **paths)


if __name__ == "__main__":
    main()
