"""Debug probe of the split-tf32 kernels (structured inputs that localise layout / precision problems)."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
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

print("=== attention tf32")
for rows, L, C in [(1, 128, 128), (1, 128, 64), (1, 128, 256), (1, 256, 128)]:
    g = torch.Generator().manual_seed(1)
    M = rows * L
    for case in ("q0", "q0_vramp", "k0", "full_small", "full"):
        qkv = torch.randn(M, 3 * C, generator=g)
        if case in ("q0", "q0_vramp"): qkv[:, :C] = 0
        if case == "q0_vramp": qkv[:, 2 * C:] = torch.arange(M, dtype=torch.float32)[:, None] + torch.arange(C)[None] * 1000.0
        if case == "k0": qkv[:, C:2 * C] = 0
        if case == "full_small": qkv[:, :2 * C] *= 0.1
        ref = att_ref(qkv, rows, L, C)
        out = torch.full((M, C), float("nan"), device=DEV)
        qk_hi, qk_lo = torch.empty(M, 2 * C, device=DEV), torch.empty(M, 2 * C, device=DEV)
        vt_hi, vt_lo = torch.empty(rows * C, L, device=DEV), torch.empty(rows * C, L, device=DEV)
        ops.attn_prep_tf32(qkv.to(DEV), qk_hi, qk_lo, vt_hi, vt_lo, rows=rows, L=L, C=C)
        ops.attention_tf32(qk_hi, qk_lo, vt_hi, vt_lo, out, rows=rows, L=L, C=C)
        torch.cuda.synchronize()
        o = out.cpu()
        print(f"rows={rows} L={L} C={C} {case}: rel {rel(o, ref):.3e} nan={int(torch.isnan(o).sum())}")
        if case == "q0_vramp" and rel(o, ref) > 1e-4:
            print("   ref row0[:8]", ref[0, :8].tolist()); print("   got row0[:8]", o[0, :8].tolist())
            print("   got row1[:4]", o[1, :4].tolist(), "row64[:4]", o[64, :4].tolist())

print("=== igemm tf32 conv precision vs K")
def nhwc(x): return x.permute(0, 2, 3, 1).contiguous()
def nchw(x): return x.permute(0, 3, 1, 2).contiguous()
for rows, H, cin, cout in [(2, 16, 64, 64), (2, 8, 256, 256), (2, 8, 512, 512), (2, 16, 128, 128)]:
    g = torch.Generator().manual_seed(2)
    a = torch.randn(rows, H, H, cin, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    ref = nhwc(F.conv2d(nchw(a.double()), w.double(), padding=1))
    wp = w.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous()
    raw = torch.empty(rows, H, H, cout, device=DEV)
    ah, al = ops.split_tf32(a.to(DEV)); wh, wl = ops.split_tf32(wp.to(DEV))
    ops.igemm((ah, al), (wh, wl), rows=rows, H=H, W=H, out_f32=raw)
    torch.cuda.synchronize()
    o = raw.cpu().double()
    ref_hh = nhwc(F.conv2d(nchw(ah.cpu().double()), wh.cpu().double().reshape(3, 3, cout, cin).permute(2, 3, 0, 1), padding=1))
    shrink = float(((o - ref) * ref.sign()).mean() / ref.abs().mean())
    print(f"K={9*cin} {cin}->{cout}: rel vs fp64 {rel(o, ref):.3e}; vs hi*hi only {rel(o, ref_hh):.3e}; hi*hi vs full {rel(ref_hh, ref):.3e}; mean signed shrink {shrink:.3e}")
    # same through the SIMT engine and the bf16 engine for reference
    raws = torch.empty(rows, H, H, cout, device=DEV)
    ops.igemm(a.to(DEV), wp.to(DEV), rows=rows, H=H, W=H, out_f32=raws)
    a16, w16 = a.to(torch.bfloat16), wp.to(torch.bfloat16)
    ref16 = nhwc(F.conv2d(nchw(a16.double()), w16.double().reshape(3, 3, cout, cin).permute(2, 3, 0, 1), padding=1))
    raw16 = torch.empty(rows, H, H, cout, device=DEV)
    ops.igemm(a16.to(DEV), w16.to(DEV), rows=rows, H=H, W=H, out_f32=raw16)
    torch.cuda.synchronize()
    o16 = raw16.cpu().double()
    print(f"      simt {rel(raws.cpu(), ref):.3e}; bf16 engine vs fp64-on-bf16-operands {rel(o16, ref16):.3e}, shrink {float(((o16 - ref16) * ref16.sign()).mean() / ref16.abs().mean()):.3e}")
