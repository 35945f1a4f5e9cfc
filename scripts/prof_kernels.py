"""Micro-driver for ncu: launches the hot kernels at bench-like shapes a few times.
usage: python scripts/prof_kernels.py [attention|conv|linear|gn|all] [rows]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spectrogramgenai_b200 import _cabi, ops
from spectrogramgenai_b200._cabi import SG_ENGINE_TC

what = sys.argv[1] if len(sys.argv) > 1 else "all"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
_cabi.require_b200(dev)
dt = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(name, fn, flops=0.0, nbytes=0.0, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {nbytes / ms / 1e6:.1f} GB/s", flush=True)


if what in ("attention", "all"):
    for (L, C) in ((4096, 64), (1024, 128), (1024, 64)):
        qkv = (torch.randn(rows * L, 3 * C, device=dev, generator=g) * 1.0).to(dt)
        out = torch.empty(rows * L, C, device=dev, dtype=dt)
        timeit(f"attention L={L} C={C} rows={rows}", lambda: ops.attention(qkv, out, rows=rows, L=L, C=C, engine=SG_ENGINE_TC),
               flops=4.0 * rows * L * L * C)
if what in ("conv", "all"):
    for (H, cin, cout) in ((64, 128, 128), (32, 256, 256), (16, 512, 512), (64, 64, 64), (8, 512, 512)):
        a = torch.randn(rows, H, H, cin, device=dev, generator=g).to(dt)
        w = (torch.randn(9, cout, cin, device=dev, generator=g) / math.sqrt(9 * cin)).to(dt)
        raw = torch.empty(rows, H, H, cout, device=dev, dtype=torch.float16)  # fp16 raw output, as the engine keeps it
        part = torch.empty(rows, ops.igemm_partials(SG_ENGINE_TC, H, H, cout), 2, device=dev)
        args = dict(rows=rows, H=H, W=H, out_act=raw, partials=part)
        _aw = (a, w)
        timeit(f"conv3x3 H={H} {cin}->{cout} rows={rows}", lambda: ops.igemm(*_aw, **args),
               flops=2.0 * rows * H * H * cin * cout * 9, nbytes=a.numel() * 2 + raw.numel() * 2)
if what in ("conv64",):  # the Cout = 64 layers (slab pipeline)
    for (H, cin, cout) in ((64, 128, 64), (64, 64, 64), (32, 128, 64)):
        a = torch.randn(rows, H, H, cin, device=dev, generator=g).to(dt)
        w = (torch.randn(9, cout, cin, device=dev, generator=g) / math.sqrt(9 * cin)).to(dt)
        raw = torch.empty(rows, H, H, cout, device=dev, dtype=torch.float16)
        part = torch.empty(rows, ops.igemm_partials(SG_ENGINE_TC, H, H, cout), 2, device=dev)
        args = dict(rows=rows, H=H, W=H, out_act=raw, partials=part)
        _aw = (a, w)
        timeit(f"conv3x3 H={H} {cin}->{cout} rows={rows}", lambda: ops.igemm(*_aw, **args),
               flops=2.0 * rows * H * H * cin * cout * 9, nbytes=a.numel() * 2 + raw.numel() * 2)
if what in ("linear", "all"):
    for (L, cin, cout) in ((4096, 64, 192), (4096, 64, 64), (1024, 128, 384)):
        M = rows * L
        a = torch.randn(M, cin, device=dev, generator=g).to(dt)
        w = (torch.randn(1, cout, cin, device=dev, generator=g) / math.sqrt(cin)).to(dt)
        b = torch.randn(cout, device=dev, generator=g)
        res = torch.randn(M, cout, device=dev, generator=g)
        o32 = torch.empty(M, cout, device=dev)
        args = dict(rows=rows, H=int(math.isqrt(L)), W=int(math.isqrt(L)), bias=b, residual=res, out_f32=o32)
        _aw = (a, w)
        timeit(f"linear M={M} {cin}->{cout}", lambda: ops.igemm(*_aw, **args), flops=2.0 * M * cin * cout,
               nbytes=a.numel() * 2 + 2 * o32.numel() * 4)
if what in ("linear16",):
    for (L, cin, cout) in ((4096, 64, 192), (1024, 128, 384)):
        M = rows * L
        a = torch.randn(M, cin, device=dev, generator=g).to(dt)
        w = (torch.randn(1, cout, cin, device=dev, generator=g) / math.sqrt(cin)).to(dt)
        b = torch.randn(cout, device=dev, generator=g)
        o16 = torch.empty(M, cout, device=dev, dtype=dt)
        args = dict(rows=rows, H=int(math.isqrt(L)), W=int(math.isqrt(L)), bias=b, out_act=o16)
        _aw = (a, w)
        timeit(f"linear16 M={M} {cin}->{cout}", lambda: ops.igemm(*_aw, **args), flops=2.0 * M * cin * cout,
               nbytes=a.numel() * 2 + o16.numel() * 2)
if what in ("gn", "all"):
    H, C = 64, 128
    raw = torch.randn(rows, H, H, C, device=dev, generator=g).half()  # fp16 raw conv output, as in the tensor-core modes
    part = torch.rand(rows, 32, 2, device=dev, generator=g) * 1000 + 1000
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    o16 = torch.empty(rows, H, H, C, device=dev, dtype=dt)
    timeit("gn_apply 64x64x128 fp16 raw -> GELU -> bf16", lambda: ops.gn_apply(raw, part, gam, bet, mode=1, out_act=o16),
           nbytes=raw.numel() * 4)
