#!/bin/bash
# full GPU test suite, decode step time + token checksums, short default bench
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; tail -4 gpurun_out/r2b_tests.log
B200W_DECODE_TIMING=1 python tools/token_checksum.py 32 5 2>&1 | grep -E "b200w_decode|tokens" | cut -c1-200 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 300 gpurun_out/r2b_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}, d.get('e2e'), d.get('roofline'), d.get('roofline_attn'), d.get('roofline_gemm'), d.get('phases_ms') or d.get('phases'))
PY
