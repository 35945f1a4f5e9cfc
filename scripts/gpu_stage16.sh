#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests (TC)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tensor_core or linear" 2>&1 | tail -3
echo "=== microbench"; python scripts/prof_kernels.py all 128 2>&1 | tail -14
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_l.json 2> gpurun_out/bench_r1_l.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_l.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items())[:6]: print(k, v)
PY
tail -3 gpurun_out/bench_r1_l.err
