import os, sys, time, torch
os.environ["B200W_DECODE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200 import frontend as FE
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like
dev = torch.device("cuda", 0)
model = B200WhisperModel("large-v3", max_batch=32, max_beam=5)
audios = [speech_like(1234 + i, 30.0) for i in range(32)]
audio_dev, offsets = FE.upload_audio(audios, dev)
kw = dict(languages=["en"] * 32, beam_size=5, temperature=0.0, word_timestamps=True)
for i in range(3):
    t = time.time(); model.transcribe_batch(None, device_audio=(audio_dev, offsets), **kw); torch.cuda.synchronize(); print("step", i, time.time() - t, flush=True)
os.environ["B200W_NO_GRAPHS"] = "1"
