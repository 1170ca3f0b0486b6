"""One step of the bench workload between cudaProfilerStart/Stop (for ncu --profile-from-start off).
usage: python tools/profile_step.py [--model large-v3] [--streams 32] [--beam 5] [--max-new 12]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import frontend as FE  # noqa: E402
from realtime_whisper_asr_b200.transcribe import B200WhisperModel  # noqa: E402
from tests.synth import speech_like  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--streams", type=int, default=32)
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--max-new", type=int, default=12)
ap.add_argument("--seconds", type=float, default=30.0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = B200WhisperModel(a.model, max_batch=a.streams, max_beam=a.beam, synthetic=True)
audios = [speech_like(1234 + i, a.seconds) for i in range(a.streams)]
audio_dev, offsets = FE.upload_audio(audios, dev)
kw = dict(languages=["en"] * a.streams, beam_size=a.beam, temperature=0.0, word_timestamps=True,
          max_new_tokens=a.max_new or None, max_rounds=1)
model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("tokens", sum(len(s.tokens) for segs, _ in out for s in segs))
