#!/bin/bash
# refresh of the bench lines (default, 50 steps, P256) after a change that does not touch the kernels' results
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo rc=$?; cut -c1-200 gpurun_out/bench_r2_final.json
timeout 900 python bench.py --steps 50 --no-cpu-baseline --no-profile > gpurun_out/bench_r2_50steps.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r2_50steps.json
timeout 600 python bench.py --size 256 --channels 1 --batch 8 --steps 5 --no-cpu-baseline > gpurun_out/bench_r2_p256.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r2_p256.json
timeout 600 python bench.py --mode fp32 --batch 64 --steps 10 --no-cpu-baseline > gpurun_out/bench_r2_fp32.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r2_fp32.json
