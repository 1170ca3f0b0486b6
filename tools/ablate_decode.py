"""Per-op cost inside the captured decode step: run the bench workload's decode with one op class
dropped at a time (B200W_ABLATE, results are garbage) and print the step time of each run.
usage (GPU box): python tools/ablate_decode.py [--steps 40]"""
import argparse
import os
import re
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--child", action="store_true")
ap.add_argument("--ops", default="", help="comma-separated op ids (default: all)")
a = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if a.child:
    import torch
    sys.path.insert(0, ROOT)
    from realtime_whisper_asr_b200 import frontend as FE
    from realtime_whisper_asr_b200.transcribe import B200WhisperModel
    from tests.synth import speech_like
    dev = torch.device("cuda", 0)
    model = B200WhisperModel("large-v3", max_batch=32, max_beam=5, synthetic=True)
    audios = [speech_like(1234 + i, 30.0) for i in range(32)]
    audio_dev, offsets = FE.upload_audio(audios, dev)
    kw = dict(languages=["en"] * 32, beam_size=5, temperature=0.0, word_timestamps=False,
              max_new_tokens=a.steps, max_rounds=1)
    for _ in range(2):
        model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
        torch.cuda.synchronize()
    sys.exit(0)

NAMES = {0: "none(1<<15)", 1: "layernorm", 2: "qkv_gemm", 3: "kv_scatter", 4: "self_attn", 5: "out_gemm", 6: "cq_gemm",
         7: "cross_attn", 8: "cout_gemm", 9: "fc1_gemm", 10: "fc2_gemm", 11: "logits_gemm", 12: "sample"}
base = None
sel = [0] + [int(x) for x in a.ops.split(',')] if a.ops else list(NAMES)
for bit, name in NAMES.items():
    if bit not in sel:
        continue
    env = dict(os.environ, B200W_DECODE_TIMING="1", B200W_ABLATE=str(1 << (15 if bit == 0 else bit)))
    out = subprocess.run([sys.executable, __file__, "--child", "--steps", str(a.steps)], env=env,
                         capture_output=True, text=True)
    ms = [float(x) for x in re.findall(r"launch\+gpu ([0-9.]+) ms", out.stderr)]
    if not ms:
        print(name, "FAILED", out.stderr[-400:])
        continue
    t = ms[-1]
    if base is None:
        base = t
        print(f"{'full step':12s} {t:8.3f} ms")
    else:
        print(f"-{name:11s} {t:8.3f} ms   => {name} costs {base - t:7.3f} ms/step")
