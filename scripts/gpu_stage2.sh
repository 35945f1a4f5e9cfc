#!/bin/bash
# Full GPU test suite, first bench line, ncu launch list of the bench command.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== bench"; timeout 1200 python bench.py > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo rc=$?; cat gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
echo "=== ncu launch list"
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 340 --csv --log-file gpurun_out/launches_r1_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu.log 2>&1
echo ncu rc=$?; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches_r1_a.csv
