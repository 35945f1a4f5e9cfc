#!/bin/bash
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 6 -o gpurun_out/prof_attention_bench -f $CMD > gpurun_out/ncu_attn_bench.log 2>&1
echo "attention full rc=$?"; ls -la gpurun_out/prof_attention_bench.ncu-rep
