#!/bin/bash
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_u.json 2> gpurun_out/bench_r1_u.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_u.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items()): print(k, v)
PY
