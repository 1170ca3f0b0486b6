"""Print a checksum of the bench workload's tokens and word times (to compare kernel variants via env switches)."""
import hashlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
beam = int(sys.argv[2]) if len(sys.argv) > 2 else 5
model = B200WhisperModel("large-v3", max_batch=n, max_beam=5, synthetic=True)
auds = [speech_like(500 + i, 30.0) for i in range(n)]
out = model.transcribe_batch(auds, languages=["en"] * n, beam_size=beam, temperature=0.0, word_timestamps=True)
toks = [t for segs, _ in out for s in segs for t in s.tokens]
words = [(round(w.start, 2), round(w.end, 2)) for segs, _ in out for s in segs for w in s.words]
print("tokens", len(toks), hashlib.md5(str(toks).encode()).hexdigest()[:12], "words", len(words),
      hashlib.md5(str(words).encode()).hexdigest()[:12], "first tokens", [segs[0].tokens[0] for segs, _ in out if segs])
