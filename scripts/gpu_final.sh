#!/bin/bash
# Round-end single-GPU evidence (no profiler): tests, smoke, bench (+ reference arm), full loop, sweeps.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "=== bench (default)"; timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo rc=$?; cut -c1-400 gpurun_out/bench_r2_final.json; tail -2 gpurun_out/bench_r2_final.err
echo "=== bench 50 steps"; timeout 900 python bench.py --steps 50 --no-cpu-baseline --no-profile > gpurun_out/bench_r2_50steps.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_50steps.json
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_reference.json
echo "=== full loop"; timeout 600 python scripts/full_loop.py > gpurun_out/full_loop_r2.json 2>&1; tail -1 gpurun_out/full_loop_r2.json
echo "=== per layer"; timeout 300 python scripts/prof_layers.py 512 64 bf16 > gpurun_out/per_layer_r2.txt 2>&1; head -1 gpurun_out/per_layer_r2.txt
echo "=== latency sweep"; timeout 900 python scripts/latency_sweep.py > gpurun_out/latency_sweep_r2.txt 2>&1; tail -3 gpurun_out/latency_sweep_r2.txt
echo "=== P256 (configs[4] geometry, 8 spectrograms on one GPU)"; timeout 600 python bench.py --size 256 --channels 1 --batch 8 --steps 5 --no-cpu-baseline > gpurun_out/bench_r2_p256.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_p256.json
timeout 300 python scripts/prof_layers.py 8 256 bf16 > gpurun_out/per_layer_p256_r2.txt 2>&1; head -1 gpurun_out/per_layer_p256_r2.txt
