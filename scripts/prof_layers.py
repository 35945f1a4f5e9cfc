"""Per-launch timing of one UNet forward at the bench geometry: every plan op with its shape, ms, TFLOP/s, GB/s.
usage: python scripts/prof_layers.py [batch=512] [size=64] [mode=bf16]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from spectrogramgenai_b200.diff_modules import Diffusion

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
c = 4 if S <= 64 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(42)
d = Diffusion(noise_steps=1000, img_size=S, num_classes=27, c_in=c, c_out=c, device=dev, compute_dtype=mode)
plan = d.model.plan(n_src=n, rows=2 * n, S=S, use_step=True)
plan.step.fill_(500)
reps = 3
acc = [0.0] * len(plan.ops)
for rep in range(reps + 1):
    evs = []
    torch.cuda.synchronize()
    for fn, a, kw in plan.ops:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*a, **kw)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if rep:
        for i, (e0, e1) in enumerate(evs):
            acc[i] += e0.elapsed_time(e1) / reps
tot = sum(acc)
print(f"total {tot:.3f} ms over {len(acc)} launches")
for (fn, a, kw), ms in zip(plan.ops, acc):
    fam, fl, by = bench.classify(fn, a, kw)
    desc = ""
    if fn.__name__ == "igemm":
        w = a[1][0] if isinstance(a[1], tuple) else a[1]
        desc = (f"{kw['H']}x{kw['W']} {w.shape[2]}->{w.shape[1]} taps={w.shape[0]} f32={kw.get('out_f32') is not None} "
                f"act={kw.get('out_act') is not None} res={kw.get('residual') is not None}")
    elif fn.__name__ == "gn_apply":
        desc = f"{tuple(a[0].shape[1:])} mode={kw.get('mode')} f32={kw.get('out_f32') is not None} act={kw.get('out_act') is not None} emb={kw.get('emb') is not None}"
    elif fn.__name__ in ("attention", "attention_tf32"):
        desc = f"L={kw['L']} C={kw['C']}"
    print(f"{fam:22s} {desc:60s} {ms:7.3f} ms {fl / ms / 1e9 if ms else 0:8.1f} TF/s {by / ms / 1e6 if ms else 0:8.1f} GB/s")
