import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import __graft_entry__ as G
G.smoke()
from realtime_whisper_asr_b200.transcribe import B200WhisperModel
from tests.synth import speech_like
m = B200WhisperModel("tiny", max_batch=4)
auds=[speech_like(i, s) for i,s in enumerate([30.0, 45.0, 3.0, 12.0, 61.0])]
t=time.time(); out=m.transcribe_batch(auds, languages=["en","de",None,"en","fr"], beam_size=5, temperature=0.0, word_timestamps=True, initial_prompt=[None,"hello there",None,None,None]); torch.cuda.synchronize(); print('batch time', time.time()-t)
for segs, info in out:
    print(info.language, round(info.duration,1), len(segs), [ (round(s.start,2), round(s.end,2), len(s.words)) for s in segs[:4]], segs[0].text[:50] if segs else None)
    w=[(x.start,x.end) for s in segs for x in s.words]
    assert all(b>=a for a,b in w)
