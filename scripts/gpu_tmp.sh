#!/bin/bash
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_t.json 2> gpurun_out/bench_r1_t.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_t.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items()): print(k, v)
PY
