"""Phase timeline of the tensor-core decode cross attention (development aid).
usage (GPU box): B200W_TRACE_CROSS=1 B200W_NO_GRAPHS=1 python tools/cross_trace.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import _lib  # noqa: E402
from realtime_whisper_asr_b200 import frontend as FE  # noqa: E402
from realtime_whisper_asr_b200.transcribe import B200WhisperModel  # noqa: E402
from tests.synth import speech_like  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
model = B200WhisperModel("large-v3", max_batch=32, max_beam=5, synthetic=True)
audios = [speech_like(1234 + i, 30.0) for i in range(32)]
audio_dev, offsets = FE.upload_audio(audios, dev)
kw = dict(languages=["en"] * 32, beam_size=5, temperature=0.0, word_timestamps=False, max_new_tokens=6, max_rounds=1)
model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
torch.cuda.synchronize()
tb = torch.zeros(296 * 6 * 8, device=dev, dtype=torch.int64)
lib.b200w_gemm_set_trace(tb.data_ptr())
model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw)
torch.cuda.synchronize()
lib.b200w_gemm_set_trace(None)
t = tb.cpu().numpy().reshape(296, 6, 8).astype(np.float64)
ok = t[:, :, 4] > 0
for it in range(6):
    m = ok[:, it]
    if not m.any():
        continue
    x = t[m, it]
    a, b, c, d = x[:, 1] - x[:, 0], x[:, 2] - x[:, 1], x[:, 3] - x[:, 2], x[:, 4] - x[:, 3]
    line = f"item {it}: ctas {m.sum():3d}  pass1(K stream) {np.median(a):7.0f}  max+pass2 {np.median(b):7.0f}  wait PV done {np.median(c):7.0f}  out {np.median(d):6.0f}"
    if it > 0:
        m2 = m & ok[:, it - 1]
        line += f"   period {np.median(t[m2, it, 0] - t[m2, it - 1, 0]):7.0f} cycles"
    print(line)
