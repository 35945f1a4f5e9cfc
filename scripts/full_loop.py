"""The complete 999-step CFG sampling loop at the bench geometry (n = 512, R64 latents, bf16 engine) through the
reference-facing API, plus the VAE decode tail: the throughput bench.py extrapolates from a few timesteps, measured
on the whole job.   usage: python scripts/full_loop.py [n] [mode]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogramgenai_b200.diff_modules import DiffusionVAE  # noqa: E402


from scripts.synth import synthetic_vqae  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
torch.manual_seed(42)
d = DiffusionVAE(noise_steps=1000, img_size=256, num_classes=27, device="cuda", vqae_state_dict=synthetic_vqae(),
                 compute_dtype=mode)  # UNet weights: the class's own random init
labels = (torch.arange(n) % 27).pin_memory()
d.sample(False, labels[:8], max_steps=2)  # module load / first-launch costs
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
img = d.sample(False, labels.to("cuda", non_blocking=True), cfg_scale=3, seed=1)
host = img.to("cpu")
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
dev = e0.elapsed_time(e1) / 1e3
hist = torch.bincount(host.flatten().long(), minlength=256)
print(json.dumps({"what": "DiffusionVAE.sample, full loop", "n": n, "mode": mode, "timesteps": 999,
                  "out_shape": list(host.shape), "device_s": round(dev, 3), "wall_s": round(wall, 3),
                  "spectrograms_per_s": round(n / dev, 3), "ms_per_timestep": round(dev / 999 * 1e3, 3),
                  "gpu_launches": d.gpu_launches, "distinct_levels": int((hist > 0).sum()),
                  "mean_level": round(float(host.float().mean()), 2)}))
