"""cProfile of the host side of one bench step (32 x 30 s, large-v3, beam 5, word timestamps): where the Python goes.
usage (GPU box): python tools/host_profile.py"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import frontend as FE
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like

n = 32
model = B200WhisperModel("large-v3", max_batch=n, max_beam=5, synthetic=True)
audios = [speech_like(1234 + i, 30.0) for i in range(n)]
kw = dict(languages=["en"] * n, beam_size=5, temperature=0.0, word_timestamps=True)
for _ in range(2):
    model.transcribe_batch(audios, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
model.transcribe_batch(audios, **kw)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
st.sort_stats("cumulative").print_stats(25)
