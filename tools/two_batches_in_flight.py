"""Two full batches in flight on ONE GPU (two model replicas, two host threads, two CUDA streams): does the
latency-bound decode of one batch overlap the encode / decode / alignment of the other?  Aggregate audio-s/s vs one
batch at a time.  usage (GPU box): python tools/two_batches_in_flight.py [steps]"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, secs = 32, 30.0
models = [B200WhisperModel("large-v3", max_batch=n, max_beam=5, synthetic=True) for _ in range(2)]
batches = [[speech_like(1000 * b + i, secs) for i in range(n)] for b in range(2)]
kw = dict(languages=["en"] * n, beam_size=5, temperature=0.0, word_timestamps=True)


def run(lane, k, out):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.device(dev), torch.cuda.stream(stream):
        for _ in range(k):
            res = models[lane].transcribe_batch(batches[lane], **kw)
        stream.synchronize()
    out[lane] = sum(len(s.tokens) for segs, _ in res for s in segs)


for lane in range(2):  # warm-up (graphs, caches)
    o = {}
    run(lane, 2, o)
torch.cuda.synchronize()
t0 = time.perf_counter()
o = {}
run(0, steps, o)
torch.cuda.synchronize()
t_seq = (time.perf_counter() - t0) / steps
print(f"one batch at a time : {t_seq * 1e3:7.1f} ms per batch  {n * secs / t_seq:7.1f} audio-s/s")
t0 = time.perf_counter()
o = {}
th = [threading.Thread(target=run, args=(lane, steps, o)) for lane in range(2)]
for t in th:
    t.start()
for t in th:
    t.join()
torch.cuda.synchronize()
t_par = (time.perf_counter() - t0) / (2 * steps)
print(f"two batches in flight: {t_par * 1e3:7.1f} ms per batch  {n * secs / t_par:7.1f} audio-s/s  (tokens {o})")
