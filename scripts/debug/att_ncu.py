"""One launch each of attention v11 and v12 at the sa6 shape (rows from argv, L = 4096, C = 64, bf16) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spectrogramgenai_b200 import ops
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256
engs = [1]
L, C = 4096, 64
g = torch.Generator(device="cuda").manual_seed(3)
qkv = torch.randn(rows * L, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(rows * L, C, device="cuda", dtype=torch.bfloat16)
for eng in engs:
    ops.attention(qkv, out, rows=rows, L=L, C=C, engine=eng)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
