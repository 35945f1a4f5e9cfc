#!/bin/bash
# Multi-GPU: weak-scaling bench through torchrun (NCCL), exactly as the driver launches it.
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
echo rc=$?; cat gpurun_out/bench_multi_$N.json | cut -c1-1800; tail -5 gpurun_out/bench_multi_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | tail -2 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/multi_gpu_check.py 2>&1 | tail -5
