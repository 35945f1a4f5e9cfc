#!/bin/bash
mkdir -p gpurun_out
echo "=== latency sweep"; timeout 900 python scripts/latency_sweep.py 64 > gpurun_out/latency_sweep_r1.txt 2> gpurun_out/latency_sweep.err; echo rc=$?; cat gpurun_out/latency_sweep_r1.txt; tail -3 gpurun_out/latency_sweep.err
echo "=== P256 (configs[4]) per-layer, 8 samples/GPU, c=1"; timeout 600 python scripts/prof_layers.py 8 256 bf16 > gpurun_out/prof_layers_p256.log 2>&1; echo rc=$?; grep -E "total|attention|fused" gpurun_out/prof_layers_p256.log | head -20; tail -3 gpurun_out/prof_layers_p256.log
