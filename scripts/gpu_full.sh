#!/bin/bash
# Full single-GPU validation: all GPU tests, smoke, bench (with CPU baseline), reference arm.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "=== bench"; timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2
echo "=== full loop"; timeout 600 python scripts/full_loop.py 2>&1 | tail -2
