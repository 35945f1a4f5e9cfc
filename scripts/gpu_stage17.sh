#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench scripts/mufu_bench.cu && /tmp/mufu_bench
for cfg in "4 0 2" "5 0 2" "5 2 2" "5 3 2"; do
  set -- $cfg
  echo "=== attention kernel tests v$1 poly8=$2 nacc=$3"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 SGB200_ATTN_NACC=$3 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention_tensor_core" 2>&1 | grep -E "passed|failed|Error|error|rel-L2 .*L=(1024|4096)" | tail -8
done
for cfg in "1 0 2" "4 0 2" "5 0 1" "5 0 2" "5 0 4" "5 1 2" "5 2 2" "5 3 2" "5 4 2"; do
  set -- $cfg
  echo "=== microbench attention v$1 poly8=$2 nacc=$3"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 SGB200_ATTN_NACC=$3 timeout 300 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
done
