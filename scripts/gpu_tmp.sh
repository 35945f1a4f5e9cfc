#!/bin/bash
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python scripts/prof_layers.py 2>&1 | grep -E "total|gn_apply.*(64, 64, 128|32, 32, 256)|^gn_apply   |tail_fused" | head
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_y.json 2> gpurun_out/bench_r1_y.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_y.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items())[:8]: print(k, v)
PY
