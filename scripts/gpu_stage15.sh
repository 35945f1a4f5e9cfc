#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_kernels.py linear 32 > gpurun_out/prof_plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 1 -c 1 -o gpurun_out/prof_linear_r1 -f python scripts/prof_kernels.py linear 32 > gpurun_out/ncu_lin.log 2>&1
echo rc=$?
python scripts/prof_kernels.py conv 64 > gpurun_out/prof_plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 1 -c 2 -o gpurun_out/prof_conv_r1b -f python scripts/prof_kernels.py conv 64 > gpurun_out/ncu_conv2.log 2>&1
echo rc=$?
cat gpurun_out/prof_plain4.log gpurun_out/prof_plain5.log
