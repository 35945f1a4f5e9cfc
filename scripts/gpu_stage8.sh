#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests (TC)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tensor_core or linear" 2>&1 | tail -3
echo "=== microbench"; python scripts/prof_kernels.py all 128 2>&1 | tail -14
echo "=== microbench attention v1"; SGB200_ATTN=1 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
echo "=== bench (attn v2)"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_f.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in d['kernels'].items(): print(k, v)
PY
echo "=== bench (attn v1)"; SGB200_ATTN=1 timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_g.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items())[:4]: print(k, v)
PY
echo "=== ncu launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 240 --csv --log-file gpurun_out/launches_r1_f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu.log 2>&1
echo ncu rc=$?
