#!/bin/bash
# Round-2 single-GPU validation: GPU tests (all failures listed), smoke, bench, fp32-engine comparison at batch 64.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -${TAIL:-80} | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
if [ "${BENCH:-1}" = "1" ]; then
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo rc=$?; cat gpurun_out/bench_r2.json; tail -3 gpurun_out/bench_r2.err
echo "=== fp32 engines at batch 64"
for m in fp32 fp32_simt; do timeout 600 python bench.py --mode $m --batch 64 --steps 5 --no-cpu-baseline > gpurun_out/bench_r2_$m.json 2> gpurun_out/bench_r2_$m.err; echo "$m rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2_$m.json")); print("$m", d["ms_per_step"], d["value"], {k:(v["ms"] if isinstance(v,dict) else v) for k,v in d["kernels"].items()})
except Exception as e: print("no json", e)
PY
tail -2 gpurun_out/bench_r2_$m.err; done
fi
