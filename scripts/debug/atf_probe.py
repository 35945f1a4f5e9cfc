"""tf32 attention: accuracy on structured inputs + timing at the sa6 / sa5 / sa1 shapes (batch 64 -> 128 rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spectrogramgenai_b200 import ops
DEV = "cuda"
def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
def att_ref(qkv, rows, L, C):
    d = C // 4
    q, k, v = qkv.double().reshape(rows, L, 3 * C).split(C, -1)
    h = lambda z: z.reshape(rows, L, 4, d).transpose(1, 2)
    att = torch.softmax(h(q) * d ** -0.5 @ h(k).transpose(-1, -2), -1) @ h(v)
    return att.transpose(1, 2).reshape(rows * L, C)
def run(qkv, rows, L, C):
    M = rows * L
    out = torch.full((M, C), float("nan"), device=DEV)
    qk_hi, qk_lo = torch.empty(M, 2 * C, device=DEV), torch.empty(M, 2 * C, device=DEV)
    vt_hi, vt_lo = torch.empty(rows * C, L, device=DEV), torch.empty(rows * C, L, device=DEV)
    ops.attn_prep_tf32(qkv.to(DEV), qk_hi, qk_lo, vt_hi, vt_lo, rows=rows, L=L, C=C)
    ops.attention_tf32(qk_hi, qk_lo, vt_hi, vt_lo, out, rows=rows, L=L, C=C)
    torch.cuda.synchronize()
    return out.cpu()
for rows, L, C in [(1, 128, 64), (1, 128, 128), (1, 128, 256), (2, 256, 64), (1, 1024, 64), (1, 4096, 64)]:
    g = torch.Generator().manual_seed(1)
    for case in ("full", "grow", "big"):
        qkv = torch.randn(rows * L, 3 * C, generator=g)
        if case == "grow": qkv[:, C:2 * C] *= torch.linspace(0.1, 6.0, rows * L)[:, None]  # later keys score higher
        if case == "big": qkv[:, :2 * C] *= 8.0
        o = run(qkv, rows, L, C)
        print(f"rows={rows} L={L} C={C} {case}: rel {rel(o, att_ref(qkv, rows, L, C)):.3e} nan={int(torch.isnan(o).sum())}", flush=True)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for rows, L, C in [(128, 4096, 64), (128, 1024, 64), (128, 1024, 128), (128, 256, 256), (512, 4096, 64)]:
    M = rows * L
    qkv = torch.randn(M, 3 * C, device=DEV)
    out = torch.empty(M, C, device=DEV)
    qk_hi, qk_lo = torch.empty(M, 2 * C, device=DEV), torch.empty(M, 2 * C, device=DEV)
    vt_hi, vt_lo = torch.empty(rows * C, L, device=DEV), torch.empty(rows * C, L, device=DEV)
    ops.attn_prep_tf32(qkv, qk_hi, qk_lo, vt_hi, vt_lo, rows=rows, L=L, C=C)
    t = min(timeit(lambda: ops.attention_tf32(qk_hi, qk_lo, vt_hi, vt_lo, out, rows=rows, L=L, C=C), reps=5) for _ in range(2))
    print(f"rows={rows} L={L} C={C}: {t:.3f} ms", flush=True)
