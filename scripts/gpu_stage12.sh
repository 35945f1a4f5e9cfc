#!/bin/bash
mkdir -p gpurun_out
echo "=== attention kernel tests (v3)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention_tensor_core" 2>&1 | grep -E "L=(256|512|1024|4096)|passed|failed|Error|error" | tail -24
for v in 3 1; do echo "=== microbench attention v$v"; SGB200_ATTN=$v SGB200_ATTN_POLY=0 python scripts/prof_kernels.py attention 128 2>&1 | tail -3; done
echo "=== model 16-bit"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "bf16 or f16 or philox" 2>&1 | grep -E "eps rel|traj|passed|failed|Error" | tail -24
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_j.json 2> gpurun_out/bench_r1_j.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_j.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items())[:5]: print(k, v)
PY
tail -3 gpurun_out/bench_r1_j.err
