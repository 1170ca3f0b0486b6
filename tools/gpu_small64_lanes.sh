#!/bin/bash
# configs[2] (small, 64 streams per GPU) with 1 and 2 concurrent lanes (sub-batches on separate engines / streams)
for l in 1 2; do
  python bench.py --config small64 --stream-seconds 30 --lanes $l 2>/dev/null | tail -1 > gpurun_out/small64_lanes$l.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/small64_lanes$l.json").read().strip().splitlines()[-1])
print("lanes", $l, round(d["value"], 1), "audio-s/s submitted,", round(d["ms_per_step"], 1), "ms per increment, unique", round(d.get("rtfx_unique_audio", 0), 1))
PY
done
