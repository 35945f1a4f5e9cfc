#!/bin/bash
echo "=== attention tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k attention 2>&1 | tail -3
python scripts/prof_kernels.py attention 128 2>&1 | tail -3
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 6 -o gpurun_out/prof_attention_bench -f $CMD > gpurun_out/ncu_attn_bench.log 2>&1
echo "attention full rc=$?"
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_v.json 2> gpurun_out/bench_r1_v.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_v.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(d['kernels'].items())[:4]: print(k, v)
PY
