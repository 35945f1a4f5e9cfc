#!/bin/bash
mkdir -p gpurun_out
echo "=== attention kernel tests"
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention_tensor_core" 2>&1 | tail -3
for cfg in "8 2" "8 3"; do
  set -- $cfg
  echo "=== microbench attention v$1 poly8=$2"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 timeout 120 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
done
