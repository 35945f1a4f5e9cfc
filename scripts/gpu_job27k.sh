#!/bin/bash
# BASELINE configs[3]: the full augmentation set -- 27 classes x 1000 samples = 27 000 spectrograms x 999 timesteps through
# the generation driver on N GPUs (VQ / decoder tail + one RGBA PNG each).  ~5 minutes on 8 B200s; PNGs (~5 GB) go to the
# roomier of /tmp and /dev/shm and are deleted afterwards.
mkdir -p gpurun_out
N=${1:-8}
S=${2:-1000}
df -k /tmp /dev/shm | tee gpurun_out/job27k_df.txt
best=$(df -k --output=avail,target /tmp /dev/shm | tail -n +2 | sort -n -r | head -1)
avail=$(echo $best | awk '{print $1}'); where=$(echo $best | awk '{print $2}')
if [ "$avail" -lt 12000000 ]; then echo "not enough scratch space ($avail KB on $where)"; exit 1; fi
timeout 1200 python scripts/gen_job.py --gpus $N --num_samples $S --work $where/sgb200_gen_job > gpurun_out/gen_job_r2_27k_${N}gpu.json 2> gpurun_out/gen_job_27k_${N}gpu.err
echo "gen_job rc=$?"; cat gpurun_out/gen_job_r2_27k_${N}gpu.json; tail -3 gpurun_out/gen_job_27k_${N}gpu.err
rm -rf $where/sgb200_gen_job
