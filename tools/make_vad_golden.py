"""Record the events of the UNCHANGED reference DynamicVADIterator (/root/reference/asr_components.py:12-78, over the
tests/harness stand-in of upstream's VADIterator and the energy stand-in of the Silero model) on seeded synthetic
streams: tests/golden/vad_gate_golden.json.  Dev container only (needs /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import harness  # noqa: E402
from tests.vad_scenarios import chunks_of, stream_audio, N_STREAMS, retunes  # noqa: E402


def main():
    assert harness.install()
    from asr_components import DynamicVADIterator
    from silero_vad_iterator import FakeSileroModel
    out = {}
    for sid in range(N_STREAMS):
        vad = DynamicVADIterator(FakeSileroModel(0.02), initial_silence_ms=500, min_silence_ms=200, max_silence_ms=1000,
                                 threshold=0.5)
        ev = []
        for tick, chunk in enumerate(chunks_of(stream_audio(sid), sid)):
            for t, ms in retunes(sid):
                if t == tick:
                    vad.set_silence_duration(ms)
            r = vad(chunk)
            ev.append(None if r is None else {k: int(v) for k, v in r.items()})
        out[str(sid)] = ev
        print(sid, "events:", [(i, e) for i, e in enumerate(ev) if e][:6], "...")
    with open(os.path.join(ROOT, "tests", "golden", "vad_gate_golden.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
