#!/bin/bash
mkdir -p gpurun_out
echo "=== conv_in/out tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_in or conv_out" 2>&1 | tail -3
for cfg in "9 2" "9 0"; do
  set -- $cfg
  echo "=== attention kernel tests v$1 poly8=$2"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention_tensor_core" 2>&1 | grep -E "passed|failed|Error|error|rel-L2 .*L=(1024|4096)|timed out" | tail -12
done
for cfg in "9 0" "9 1" "9 2" "9 3" "8 2"; do
  set -- $cfg
  echo "=== microbench attention v$1 poly8=$2"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 timeout 120 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
done
SGB200_ATTN=9 SGB200_ATTN_POLY8=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc9 -s 1 -c 1 -o gpurun_out/prof_attn_v9p2 -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn9.log 2>&1
echo ncu rc=$?
