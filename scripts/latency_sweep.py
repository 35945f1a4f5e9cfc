"""BASELINE configs[1]: UNet_conditional single denoising step, latency sweep over batch 1..256, fp32 and 16-bit engines,
one B200.  One step = cond + uncond forward (2n rows) + CFG lerp + posterior update, replayed from a CUDA graph.
Also prints the per-step eps parity of each engine against the CUDA-core fp32 engine at the same inputs (which, like the
others, is pinned to the reference by tests/test_gpu_model.py: n <= 5 from golden vectors, n = 64 against the oracle).
Engines: fp32_simt = CUDA-core comparator (n <= 64), fp32 = split-TF32 tensor cores (fp32-accurate), bf16 / f16 = tcgen05.
usage: python scripts/latency_sweep.py [size=64] > profiles/latency_sweep_r2.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spectrogramgenai_b200 import ops
from spectrogramgenai_b200.diff_modules import Diffusion

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
print(f"# UNet_conditional(c_in=4, c_out=4, num_classes=27) at [4,{S},{S}], CFG step = 2n rows; CUDA-graph replay, median of 20")
print(f"{'batch':>6s} {'mode':>9s} {'ms/step':>9s} {'steps/s':>9s} {'spectrograms/s (999 steps)':>27s} {'eps rel-L2 vs fp32_simt':>26s}")
ref_eps = {}
for mode in ("fp32_simt", "fp32", "bf16", "f16"):
    torch.manual_seed(42)
    d = Diffusion(noise_steps=1000, img_size=S, num_classes=27, c_in=4, c_out=4, device=dev, compute_dtype=mode)
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        if mode == "fp32_simt" and n > 64:
            continue  # the CUDA-core engine is the comparator, not a throughput path
        plan = d.model.plan(n_src=n, rows=2 * n, S=S, use_step=True)
        g = torch.Generator(device="cpu").manual_seed(123)
        x0 = torch.randn((n, 4, S, S), generator=g).to(dev)
        plan.x_in.copy_(x0)
        plan.y.fill_(-1)
        plan.y[:n].copy_((torch.arange(n) % 27).to(dev))
        plan.step.fill_(500)
        plan.run()
        torch.cuda.synchronize()
        eps = plan.eps.clone()
        if mode == "fp32_simt":
            ref_eps[n] = eps
        err = float((eps - ref_eps[n]).norm() / ref_eps[n].norm()) if n in ref_eps else float("nan")

        def one_step():
            plan.run()
            ops.cfg_update(plan.x_in, plan.eps, d._coef, plan.step, cfg_scale=3.0, seed=1, sample_base=0)

        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            one_step()
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            plan.x_in.copy_(x0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        print(f"{n:6d} {mode:>9s} {ms:9.3f} {1e3 / ms:9.1f} {n / (ms * 1e-3 * 999):27.3f} {err:26.3e}", flush=True)
        d.model.release_plans()
        del plan, gr
        torch.cuda.empty_cache()
