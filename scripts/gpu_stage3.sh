#!/bin/bash
mkdir -p gpurun_out
echo "=== attention tc kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention_tensor_core" 2>&1 | tail -40
echo "=== model 16-bit"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "bf16 or f16 or philox" 2>&1 | grep -v "^tap" | tail -30
echo "=== bench"; timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; echo rc=$?; cat gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
