#!/bin/bash
mkdir -p gpurun_out
echo "=== conv tests (igemm v3 halo)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_tensor_core" 2>&1 | tail -3
for v in 3 2; do echo "=== microbench conv igemm v$v"; SGB200_IGEMM=$v python scripts/prof_kernels.py conv 128 2>&1 | tail -5; done
echo "=== model 16-bit"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "bf16 or f16" 2>&1 | grep -E "eps rel|traj|passed|failed|Error" | tail -20
SGB200_ATTN=3 python scripts/prof_kernels.py attention 32 > gpurun_out/prof_plain3.log 2>&1 && \
SGB200_ATTN=3 ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 1 -c 1 -o gpurun_out/prof_attn_v3 -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn3.log 2>&1
echo ncu rc=$?
