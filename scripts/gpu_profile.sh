#!/bin/bash
# Round-end evidence, ONE ncu command per gpurun call (each follows a plain run of the same command that exited 0):
#   bash scripts/gpu_profile.sh launches   -> gpurun_out/ncu_launches_r2.csv   (launch list of the bench command)
#   bash scripts/gpu_profile.sh attention  -> gpurun_out/prof_attention_bench.ncu-rep (six attention launches, --set full)
#   bash scripts/gpu_profile.sh fused      -> gpurun_out/prof_membound.ncu-rep        (fused SelfAttention + Up-block kernels)
#   bash scripts/gpu_profile.sh conv|gn    -> gpurun_out/prof_{conv,gn}_bench.ncu-rep (scripts/prof_kernels.py shapes)
# Summaries for profiles/: python scripts/ncu_summary.py gpurun_out/<file>.ncu-rep
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
case "${1:-launches}" in
  launches)
    $CMD > gpurun_out/plain_bench.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_r2.csv $CMD > gpurun_out/ncu_launches.log 2>&1 ;;
  attention)
    $CMD > gpurun_out/plain_bench.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 6 -o gpurun_out/prof_attention_bench -f $CMD > gpurun_out/ncu_attn_bench.log 2>&1 ;;
  fused)
    $CMD > gpurun_out/plain_bench.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k "regex:gn_apply_vcat|upsample_cat|attn_tail|ln_inproj" -c 14 -o gpurun_out/prof_membound -f $CMD > gpurun_out/ncu_membound.log 2>&1 ;;
  conv)
    python scripts/prof_kernels.py conv 1024 > gpurun_out/plain_conv.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 5 -c 5 -o gpurun_out/prof_conv_bench -f python scripts/prof_kernels.py conv 1024 > gpurun_out/ncu_conv_bench.log 2>&1 ;;
  gn)
    python scripts/prof_kernels.py gn 1024 > gpurun_out/plain_gn.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 1 -c 1 -o gpurun_out/prof_gn_bench -f python scripts/prof_kernels.py gn 1024 > gpurun_out/ncu_gn_bench.log 2>&1 ;;
esac
echo "ncu rc=$?"
