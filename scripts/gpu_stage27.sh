#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_bench scripts/pipe_bench.cu && /tmp/pipe_bench
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc8 -s 1 -c 1 -o gpurun_out/prof_attn_v8b -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn8b.log 2>&1
echo ncu rc=$?
