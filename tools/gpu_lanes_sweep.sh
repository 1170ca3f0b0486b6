#!/bin/bash
# Bench throughput for several lane counts (concurrent sub-batches per GPU).
for l in 1 2 4; do
  echo "lanes=$l"
  python bench.py --steps 3 --warmup 2 --no-side-legs --no-cpu-baseline --lanes $l 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    try: j=json.loads(ln)
    except Exception: continue
    print('value %.1f e2e %.1f ms_per_step %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']))"
done
