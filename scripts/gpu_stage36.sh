#!/bin/bash
mkdir -p gpurun_out
echo "=== default bench"; ( time timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -4; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['roofline'], d['cpu_baseline'])
PY
echo "=== reference arm"; ( time timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -4; cat gpurun_out/bench_ref.json | cut -c1-600
