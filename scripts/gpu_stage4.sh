#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_kernels.py all 128 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_attn_r1 -f python scripts/prof_kernels.py attention 32 > gpurun_out/ncu_attn.log 2>&1
echo rc=$?; cat gpurun_out/prof_plain.log
ncu --set full --clock-control none --import-source on -k regex:igemm_tc -s 1 -c 1 -o gpurun_out/prof_conv_r1 -f python scripts/prof_kernels.py conv 32 > gpurun_out/ncu_conv.log 2>&1
echo rc=$?
