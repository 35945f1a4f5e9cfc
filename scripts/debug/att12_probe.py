"""v12 attention probe: parity on the kernel-test shapes (incl. the overflow / growing-score cases) and timing vs v11."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spectrogramgenai_b200 import ops
DEV = "cuda"

def ref(qkv, rows, L, C):
    d = C // 4
    q, k, v = qkv.double().reshape(rows, L, 3 * C).split(C, -1)
    h = lambda z: z.reshape(rows, L, 4, d).transpose(1, 2)
    att = torch.softmax(h(q) * d ** -0.5 @ h(k).transpose(-1, -2), -1) @ h(v)
    return att.transpose(1, 2).reshape(rows * L, C)

def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm())

def run(qkv, rows, L, C, eng):
    out = torch.full((rows * L, C), float("nan"), device=DEV, dtype=qkv.dtype)
    ops.attention(qkv.to(DEV), out, rows=rows, L=L, C=C, engine=eng)
    torch.cuda.synchronize()
    return out.cpu()

for dt in (torch.bfloat16, torch.float16):
    for rows, L in [(1, 128), (2, 256), (1, 1024), (2, 4096)]:
        C = 64
        qkv = (torch.randn(rows * L, 3 * C, generator=torch.Generator().manual_seed(12)) * 1.5).to(dt)
        print(f"{dt} rows={rows} L={L}: rel {rel(run(qkv, rows, L, C, 1), ref(qkv.float(), rows, L, C)):.3e}")
print("=== timing (bf16)")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for rows, L in [(128, 4096), (256, 4096), (1024, 4096), (1024, 1024)]:
    C = 64
    g = torch.Generator(device=DEV).manual_seed(3)
    qkv = torch.empty(rows * L, 3 * C, device=DEV, dtype=torch.bfloat16)
    for r0 in range(0, rows, 128):
        qkv[r0 * L:(r0 + 128) * L] = torch.randn(128 * L, 3 * C, device=DEV, generator=g).to(torch.bfloat16)
    out = torch.empty(rows * L, C, device=DEV, dtype=torch.bfloat16)
    res = {}
    for eng in (1,):
        res[eng] = min(timeit(lambda: ops.attention(qkv, out, rows=rows, L=L, C=C, engine=eng), reps=5 if rows >= 1024 else 10) for _ in range(3))
    print(f"rows={rows} L={L}: " + "  ".join(f"eng{e} {t:.3f} ms" for e, t in res.items()), flush=True)
