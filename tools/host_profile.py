"""cProfile of one bench step (32 streams, large-v3): where does the host side go? (development aid)"""
import cProfile, os, pstats, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import frontend as FE
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like
model = B200WhisperModel("large-v3", max_batch=32, max_beam=5)
audios = [speech_like(1234 + i, 30.0) for i in range(32)]
kw = dict(languages=["en"] * 32, beam_size=5, temperature=0.0, word_timestamps=True)
for _ in range(2):
    model.transcribe_batch(audios, **kw)
pr = cProfile.Profile()
pr.enable()
model.transcribe_batch(audios, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
