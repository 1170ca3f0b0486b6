"""Where does one live-style transcribe() call of Whisper base go?  (development aid, GPU box)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200.backend import B200WhisperASR
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like

model = B200WhisperModel("base", device="cuda", device_index=0, max_batch=1, max_beam=5, synthetic=True)
asr = B200WhisperASR("en", model=model, transcribe_kwargs={"beam_size": 5, "temperature": 0.0})
audio = speech_like(4321, 30.0)
for i in range(1, 20):
    buf = audio[max(0, i - 15) * 16000: i * 16000]
    if i >= 17:
        model.profile_phases = True
        model.last_timings = {}
        os.environ["B200W_DECODE_TIMING"] = "1"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = asr.transcribe(buf, init_prompt="previous words")
    t1 = time.perf_counter()
    w = asr.ts_words(res)
    t2 = time.perf_counter()
    if i >= 17:
        print(f"call {i}: transcribe {1000 * (t1 - t0):.1f} ms, ts_words {1000 * (t2 - t1):.2f} ms, words {len(w)}, "
              f"tokens {sum(len(s.tokens) for s in res)}; phases(ms) "
              + ", ".join(f"{k}={1000 * v:.1f}" for k, v in model.last_timings.items() if k != '_'), flush=True)
import cProfile, pstats
model.profile_phases = False
os.environ.pop("B200W_DECODE_TIMING", None)
pr = cProfile.Profile()
pr.enable()
for i in range(20, 25):
    buf = audio[max(0, i - 15) * 16000: i * 16000]
    asr.ts_words(asr.transcribe(buf, init_prompt="previous words"))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
