"""Record what the UNCHANGED reference streaming policy commits on scripted hypothesis sequences:
tests/golden/streaming_policy_golden.json.  Runs in the dev container only (needs /root/reference); the stand-ins of
the un-vendored upstream modules come from tests/harness.

For every scenario of tests/streaming_scenarios.py one /root/reference/enhanced_asr_processor.py
EnhancedOnlineASRProcessor is driven by a scripted backend (ScriptedModel: a deterministic function of the audio it
is handed -- the samples encode their absolute time -- that emits a slowly stabilising word sequence, like streaming
ASR does); recorded per iteration: the prompt the backend received, the returned (beg, end, text), the buffer offset
and length after the iteration."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import harness  # noqa: E402
from tests.streaming_scenarios import SCENARIOS, ScriptedModel, chunks_of  # noqa: E402


class _ASR:
    """ASRBase-shaped backend over the scripted model (what CustomFasterWhisperASR is over faster-whisper)."""
    sep = ""

    def __init__(self, model, language):
        self.model, self.language, self.prompts = model, language, []

    def transcribe(self, audio, init_prompt=""):
        self.prompts.append(init_prompt)
        segs, _ = self.model.transcribe_batch([audio], languages=[self.language], initial_prompts=[init_prompt])[0]
        return list(segs)

    def ts_words(self, segments):
        return [(w.start, w.end, w.word) for s in segments for w in s.words]

    def segments_end_ts(self, segments):
        return [s.end for s in segments]


def main():
    assert harness.install(), "the reference tree is needed to record the golden"
    import unittest.mock as mock
    with mock.patch("os.path.exists", lambda p, _e=os.path.exists: True if "whisper_streaming-main" in str(p) else _e(p)):
        import enhanced_asr_processor as E
    out = {}
    for name, sc in SCENARIOS.items():
        model = ScriptedModel(sc["seed"], fail_at=sc.get("fail_at"))
        asr = _ASR(model, "en")
        proc = E.EnhancedOnlineASRProcessor(asr, buffer_trimming=("segment", sc["trim"]), agreement_n=sc["n"],
                                            enable_dynamic_buffer=sc["dynamic"], logfile=open(os.devnull, "w"))
        rec = []
        for chunk in chunks_of(sc):
            proc.insert_audio_chunk(chunk)
            n_before = len(asr.prompts)
            res = proc.process_iter()
            rec.append(dict(prompt=asr.prompts[-1] if len(asr.prompts) > n_before else None, out=list(res),
                            offset=proc.buffer_time_offset, buffer_len=len(proc.audio_buffer),
                            trimming_sec=proc.buffer_trimming_sec))
        rec.append(dict(finish=list(proc.finish()), offset=proc.buffer_time_offset))
        out[name] = rec
        print(name, len(rec) - 1, "iterations,", sum(1 for r in rec[:-1] if r["out"][2]), "with committed text, final offset",
              rec[-1]["offset"])
    with open(os.path.join(ROOT, "tests", "golden", "streaming_policy_golden.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
