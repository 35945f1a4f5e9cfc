"""v8 (d = 32 / 64) attention: timing at the sa1 / sa2 / sa4 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spectrogramgenai_b200 import ops
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for rows, L, C in [(1024, 1024, 128), (1024, 256, 256), (1024, 256, 128), (16, 16384, 128)]:
    qkv = torch.randn(rows * L, 3 * C, device="cuda").to(torch.bfloat16)
    out = torch.empty(rows * L, C, device="cuda", dtype=torch.bfloat16)
    t = min(timeit(lambda: ops.attention(qkv, out, rows=rows, L=L, C=C, engine=1)) for _ in range(3))
    print(f"rows={rows} L={L} C={C}: {t:.3f} ms", flush=True)
