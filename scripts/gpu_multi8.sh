#!/bin/bash
# 8-GPU evidence: the north-star job through the generation driver, then the weak-scaling bench line.
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python scripts/gen_job.py --gpus $N --num_samples 37 > gpurun_out/gen_job_r2_${N}gpu.json 2> gpurun_out/gen_job_${N}gpu.err; echo "gen_job rc=$?"; cat gpurun_out/gen_job_r2_${N}gpu.json; tail -3 gpurun_out/gen_job_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; echo "bench rc=$?"; cat gpurun_out/bench_r2_${N}gpu.json | cut -c1-1500; tail -2 gpurun_out/bench_r2_${N}gpu.err
