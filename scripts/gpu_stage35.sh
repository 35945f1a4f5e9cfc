#!/bin/bash
python scripts/prof_kernels.py linear16 256 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 2 -c 1 -o gpurun_out/prof_linear16 -f python scripts/prof_kernels.py linear16 256 > gpurun_out/ncu_lin16.log 2>&1
echo ncu rc=$?
