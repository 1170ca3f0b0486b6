"""Phase timeline of one fused decode-step program (dstep.cu): per phase, when the vector warps finished their work
and when they left the grid barrier, averaged over CTAs.  usage (GPU box): B200W_TRACE_DSTEP=<key> python tools/dstep_trace.py
key: 0 = K0, 1 + 2 l = program A of layer l, 2 + 2 l = program B of layer l."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import _lib
from realtime_whisper_asr_b200 import frontend as FE
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like

lib = _lib.load()
dev = torch.device("cuda", 0)
buf = torch.zeros((148, 24), dtype=torch.int64, device=dev)
lib.b200w_gemm_set_trace(buf.data_ptr())
model = B200WhisperModel("large-v3", max_batch=32, max_beam=5, synthetic=True)
audios = [speech_like(1234 + i, 30.0) for i in range(32)]
audio_dev, offsets = FE.upload_audio(audios, dev)
kw = dict(languages=["en"] * 32, beam_size=5, temperature=0.0, word_timestamps=False, max_new_tokens=12, max_rounds=1)
model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
print("kernel start spread (us): %.2f .. %.2f; pdl wait passed at %.2f (mean)" % ((t[:, 0].min() - t0) / 1e3, (t[:, 0].max() - t0) / 1e3, (t[:, 1].mean() - t0) / 1e3))
prev = t[:, 1].mean()
for p in range(8):
    a, b = t[:, 2 + 2 * p], t[:, 3 + 2 * p]
    if a.max() == 0:
        break
    print("phase %d: work done mean %.2f max %.2f us after kernel start (%.2f us after previous barrier)" %
          (p, (a.mean() - t0) / 1e3, (a.max() - t0) / 1e3, (a.max() - prev) / 1e3), end="")
    if b.max() > 0:
        print("; barrier left at %.2f (mean), barrier cost after last arrival %.2f us" % ((b.mean() - t0) / 1e3, (b.mean() - a.max()) / 1e3))
        prev = b.mean()
    else:
        print()
