#!/bin/bash
mkdir -p gpurun_out
{
  timeout 300 python scripts/debug/att12_probe.py 2>&1 | tail -14
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" 2>&1 | tail -2
  timeout 300 python bench.py --steps 6 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_wg.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_wg.json')); print(d['ms_per_step'], d['value'], d['roofline']['launch_ms'], d['clocks'])"
} > gpurun_out/att12_wg.txt 2>&1
cat gpurun_out/att12_wg.txt
