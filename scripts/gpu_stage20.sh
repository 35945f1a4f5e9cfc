#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 2" "8 0"; do
  set -- $cfg
  echo "=== attention kernel tests v$1 poly8=$2"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention_tensor_core" 2>&1 | grep -E "passed|failed|Error|error|rel-L2 .*L=(1024|4096)|timed out" | tail -12
done
for cfg in "8 0" "8 1" "8 2" "8 3" "5 2"; do
  set -- $cfg
  echo "=== microbench attention v$1 poly8=$2"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 timeout 300 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
done
SGB200_ATTN=8 SGB200_ATTN_POLY8=2 ncu --set full --clock-control none --import-source on -k regex:attention_tc8 -s 1 -c 1 -o gpurun_out/prof_attn_v8p2 -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn8.log 2>&1
echo ncu rc=$?
