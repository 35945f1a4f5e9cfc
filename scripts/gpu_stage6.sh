#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_in or groupnorm" 2>&1 | tail -4
python scripts/prof_kernels.py attention 32 > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_attn_r1b -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn2.log 2>&1
echo rc=$?
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_d.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in d['kernels'].items(): print(k, v)
PY
