"""Memory-bound kernels at bench shapes (rows = 2n = 1024 by default): achieved GB/s against algorithmic bytes.
usage: python scripts/prof_mem.py [rows]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spectrogramgenai_b200 import _cabi, ops
from spectrogramgenai_b200._cabi import SG_ENGINE_TC

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
_cabi.require_b200(dev)
dt = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(name, fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:58s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)


for (H, C) in ((64, 64), (32, 128), (64, 128)):
    raw = torch.randn(rows, H, H, C, device=dev, generator=g).half()  # fp16 raw conv output (tensor-core modes)
    res = torch.randn(rows, H, H, C, device=dev, generator=g)
    P = H * H // 128
    part = torch.rand(rows, P, 2, device=dev, generator=g) * 1000 + 1000
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    o16 = torch.empty(rows, H, H, C, device=dev, dtype=dt)
    o32 = torch.empty(rows, H, H, C, device=dev)
    emb = torch.randn(rows, C, device=dev, generator=g)
    n = raw.numel()
    timeit(f"gn_apply {H}x{H}x{C} mode1 (GELU) -> bf16", lambda: ops.gn_apply(raw, part, gam, bet, mode=1, out_act=o16), n * 4)
    timeit(f"gn_apply {H}x{H}x{C} mode0 -> bf16", lambda: ops.gn_apply(raw, part, gam, bet, mode=0, out_act=o16), n * 4)
    timeit(f"gn_apply {H}x{H}x{C} mode0 +emb -> f32", lambda: ops.gn_apply(raw, part, gam, bet, mode=0, emb=emb, out_f32=o32), n * 6)
    timeit(f"gn_apply {H}x{H}x{C} mode2 (res+GELU) -> bf16", lambda: ops.gn_apply(raw, part, gam, bet, mode=2, residual=res, out_act=o16), n * 8)
    timeit(f"gn_apply {H}x{H}x{C} mode0 -> f32+bf16", lambda: ops.gn_apply(raw, part, gam, bet, mode=0, out_f32=o32, out_act=o16), n * 8)
    del raw, res, o16, o32
for (L, C) in ((4096, 64), (1024, 128), (1024, 64)):
    M = rows * L
    x = torch.randn(M, C, device=dev, generator=g)
    o16 = torch.empty(M, C, device=dev, dtype=dt)
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    timeit(f"layernorm M={M} C={C} -> bf16", lambda: ops.layernorm(x, gam, bet, o16), M * C * 6)
    del x, o16
for (L, cin, cout, resid, gelu, o32f, o16f) in ((4096, 64, 192, False, False, False, True), (4096, 64, 64, True, False, True, False),
                                                 (4096, 64, 64, False, True, False, True), (4096, 64, 64, True, False, True, False),
                                                 (1024, 128, 384, False, False, False, True), (1024, 128, 128, True, False, True, False)):
    M = rows * L
    a = torch.randn(M, cin, device=dev, generator=g).to(dt)
    w = (torch.randn(1, cout, cin, device=dev, generator=g) / math.sqrt(cin)).to(dt)
    b = torch.randn(cout, device=dev, generator=g)
    res = torch.randn(M, cout, device=dev, generator=g) if resid else None
    o32 = torch.empty(M, cout, device=dev) if o32f else None
    o16 = torch.empty(M, cout, device=dev, dtype=dt) if o16f else None
    H = int(math.isqrt(L))
    args = dict(rows=rows, H=H, W=H, bias=b, residual=res, gelu=gelu, out_f32=o32, out_act=o16)
    _aw = (a, w)
    nb = a.numel() * 2 + (res.numel() * 4 if resid else 0) + (o32.numel() * 4 if o32f else 0) + (o16.numel() * 2 if o16f else 0)
    timeit(f"linear M={M} {cin}->{cout} res={resid} gelu={gelu} f32={o32f} bf16={o16f}", lambda: ops.igemm(*_aw, **args), nb)
    del a, res, o32, o16
x = torch.randn(rows, 32, 32, 64, device=dev, generator=g)
skip = torch.randn(rows, 64, 64, 64, device=dev, generator=g)
o32 = torch.empty(rows, 64, 64, 128, device=dev)
o16 = torch.empty(rows, 64, 64, 128, device=dev, dtype=dt)
timeit("upsample_cat 32->64, 64+64 ch -> f32+bf16", lambda: ops.upsample_cat(x, skip, out_f32=o32, out_act=o16),
       x.numel() * 4 + skip.numel() * 4 + o32.numel() * 6)
