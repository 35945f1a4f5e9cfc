#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench scripts/mufu_bench.cu && /tmp/mufu_bench
SGB200_ATTN=5 SGB200_ATTN_POLY8=2 ncu --set full --clock-control none --import-source on -k regex:attention_tc5 -s 1 -c 1 -o gpurun_out/prof_attn_v5p2 -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn5.log 2>&1
echo ncu rc=$?
SGB200_ATTN=5 SGB200_ATTN_POLY8=0 ncu --set full --clock-control none --import-source on -k regex:attention_tc5 -s 1 -c 1 -o gpurun_out/prof_attn_v5p0 -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn5b.log 2>&1
echo ncu rc=$?
