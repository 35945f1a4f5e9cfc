#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 2" "8 0" "8 3"; do
  set -- $cfg
  echo "=== attention kernel tests v$1 poly8=$2"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention_tensor_core" 2>&1 | grep -E "passed|failed|Error|error|timed out|growing" | tail -12
done
for cfg in "8 0" "8 1" "8 2" "8 3"; do
  set -- $cfg
  echo "=== microbench attention v$1 poly8=$2"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 timeout 120 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
done
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_n.json 2> gpurun_out/bench_r1_n.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_n.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items()): print(k, v)
PY
tail -3 gpurun_out/bench_r1_n.err
