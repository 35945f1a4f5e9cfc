#!/bin/bash
# Round-end evidence: launch list of the bench command, `ncu --set full` of the dominant kernels at bench shapes.
# (one gpurun call; every ncu command follows a plain run of the same command that exited 0)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ncu_launches_r1_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 6 -o gpurun_out/prof_attention_bench -f $CMD > gpurun_out/ncu_attn_bench.log 2>&1
echo "attention full rc=$?"
python scripts/prof_kernels.py conv 1024 > gpurun_out/plain_conv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 5 -c 5 -o gpurun_out/prof_conv_bench -f python scripts/prof_kernels.py conv 1024 > gpurun_out/ncu_conv_bench.log 2>&1
echo "conv full rc=$?"
python scripts/prof_kernels.py gn 1024 > gpurun_out/plain_gn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 1 -c 1 -o gpurun_out/prof_gn_bench -f python scripts/prof_kernels.py gn 1024 > gpurun_out/ncu_gn_bench.log 2>&1
echo "gn full rc=$?"
tail -3 gpurun_out/plain_conv.log gpurun_out/plain_gn.log
