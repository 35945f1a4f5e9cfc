#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
echo "=== model tests"; timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_s.json 2> gpurun_out/bench_r1_s.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_s.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items()): print(k, v)
PY
