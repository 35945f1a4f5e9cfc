"""Per-kernel parity of the C-ABI CUDA kernels against the CPU oracle / torch fp32-fp64 restatements.

All tests call through spectrogramgenai_b200.ops, i.e. through the ctypes binding of libsgb200.so.
Tolerances: bit-exact for the sampler arithmetic and the uint8 tail; fp32 kernels within 2e-5 rel-L2 of an
fp64 evaluation; tensor-core kernels within 2e-5 of an fp64 evaluation ON THE SAME 16-bit-rounded operands
(so the test isolates kernel correctness from operand rounding, which the end-to-end tests bound).
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ddpm_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from spectrogramgenai_b200 import _cabi, ops as _ops

    _cabi.require_b200(torch.device("cuda", torch.cuda.current_device()))
    return _ops


def gen(seed):
    return torch.Generator().manual_seed(seed)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def pack_conv(w, dtype=torch.float32):
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous().to(dtype)


# ------------------------------------------------------------------------------------------ sampler
@pytest.mark.parametrize("cfg", [3.0, 0.3, 0.0])
@pytest.mark.parametrize("i", [999, 500, 2, 1])
def test_cfg_update_bit_exact(ops, cfg, i):
    n, c, S, T = 3, 4, 16, 1000
    g = gen(i)
    x = torch.randn(n, c, S, S, generator=g)
    eps = torch.randn(2 * n, c, S, S, generator=g)
    noise = torch.randn(T - 1, n, c, S, S, generator=g)
    beta, alpha, ah = O.noise_schedule(T)
    c1, c2, c3 = O.posterior_coefficients(beta, alpha, ah)
    e = O.cfg_combine(eps[:n], eps[n:], cfg) if cfg > 0 else eps[:n]
    z = noise[T - i] if i > 1 else torch.zeros_like(x)
    ref = O.posterior_update(x, e, c1[i], c2[i], c3[i], z)
    coef = torch.stack([c1, c2, c3], 1).contiguous().to(DEV)
    xd = x.to(DEV)
    step = torch.tensor([i], dtype=torch.int32, device=DEV)
    ops.cfg_update(xd, eps.to(DEV), coef, step, cfg_scale=cfg, noise=noise.to(DEV))
    assert torch.equal(xd.cpu(), ref)
    ops.step_advance(step)
    assert int(step.item()) == i - 1


@pytest.mark.parametrize("count", [0, 1, 3, 4, 1023, 4 * 16 * 16 * 5])
def test_to_uint8_bit_exact(ops, count):
    g = gen(count)
    x = torch.randn(count, generator=g) * 0.8
    if count >= 4:
        x[:4] = torch.tensor([-1.0, 1.0, -7.0, 7.0])
    if count > 8:
        x[4:8] = torch.tensor([0.0, 0.999999, -0.999999, 0.5])
    out = torch.zeros(max(count, 1), dtype=torch.uint8, device=DEV)
    ops.to_uint8(x.to(DEV), out) if count else None
    assert torch.equal(out.cpu()[:count], O.to_uint8(x))


def test_philox_normal_statistics_and_sharding_invariance(ops):
    n, E = 8, 4 * 64 * 64
    a = torch.empty(n, E, device=DEV)
    ops.philox_normal(a, seed=1234, sample_base=0, step_tag=1000)
    v = a.double().cpu()
    assert abs(float(v.mean())) < 5e-3 and abs(float(v.std()) - 1) < 5e-3
    assert abs(float((v ** 3).mean())) < 2e-2 and abs(float((v ** 4).mean()) - 3) < 5e-2
    assert torch.isfinite(a).all()
    # the stream is keyed by the GLOBAL sample index: two half-batches == one batch
    b = torch.empty(4, E, device=DEV)
    ops.philox_normal(b, seed=1234, sample_base=4, step_tag=1000)
    assert torch.equal(b, a[4:])
    # different step / seed -> different numbers
    ops.philox_normal(b, seed=1234, sample_base=4, step_tag=999)
    assert not torch.equal(b, a[4:])
    # cfg_update's in-kernel noise == philox_normal of the same (seed, sample, step)
    T = 10
    coef = torch.tensor([[1.0, 0.0, 1.0]] * T, device=DEV)
    x = torch.zeros(4, E, device=DEV)
    step = torch.tensor([7], dtype=torch.int32, device=DEV)
    ops.cfg_update(x.view(4, 4, 64, 64), torch.zeros(8, 4, 64, 64, device=DEV), coef, step, cfg_scale=3.0, seed=1234,
                   sample_base=4)
    ops.philox_normal(b, seed=1234, sample_base=4, step_tag=7)
    assert torch.equal(x, b)


# ------------------------------------------------------------------------------------------ pool / upsample
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (3, 2, 2, 256), (1, 64, 64, 64)])
def test_maxpool2(ops, shape):
    x = torch.randn(shape, generator=gen(1))
    ref = nhwc(F.max_pool2d(nchw(x), 2))
    o32 = torch.empty(ref.shape, device=DEV)
    o16 = torch.empty(ref.shape, device=DEV, dtype=torch.bfloat16)
    ops.maxpool2(x.to(DEV), out_f32=o32, out_act=o16)
    assert torch.equal(o32.cpu(), ref)
    assert torch.equal(o16.cpu(), ref.to(torch.bfloat16))


@pytest.mark.parametrize("h,cx,cs", [(2, 256, 256), (8, 128, 128), (32, 64, 64), (1, 64, 64), (4, 64, 128), (4, 12, 4), (16, 8, 8)])
def test_upsample_cat(ops, h, cx, cs):
    rows = 2
    x = torch.randn(rows, h, h, cx, generator=gen(2))
    skip = torch.randn(rows, 2 * h, 2 * h, cs, generator=gen(3))
    up = F.interpolate(nchw(x).double(), scale_factor=2, mode="bilinear", align_corners=True)
    ref = nhwc(torch.cat([nchw(skip).double(), up], 1))
    o32 = torch.empty(ref.shape, device=DEV)
    o16 = torch.empty(ref.shape, device=DEV, dtype=torch.float16)
    ops.upsample_cat(x.to(DEV), skip.to(DEV), out_f32=o32, out_act=o16)
    # fp32 source coordinates (dst * (in-1)/(out-1)) as torch computes them: compare with torch fp32 tightly,
    # with the fp64 evaluation loosely
    ref32 = nhwc(torch.cat([nchw(skip), F.interpolate(nchw(x), scale_factor=2, mode="bilinear", align_corners=True)], 1))
    assert float((o32.cpu() - ref32).abs().max()) < 2e-6
    assert float((o32.cpu().double() - ref).abs().max()) < 3e-5
    assert float((o16.cpu().double() - ref).abs().max()) < 4e-3


# ------------------------------------------------------------------------------------------ time embedding
def test_time_embed(ops):
    rows, ncls = 11, 27
    g = gen(4)
    t = torch.tensor([999, 998, 500, 20, 2, 1, 0, 7, 300, 640, 999]).float()
    y = torch.randint(0, ncls, (rows,), generator=g)
    y[3] = -1
    y[10] = -1
    label = torch.randn(ncls, 256, generator=g)
    w = torch.randn(896, 256, generator=g) / 16
    b = torch.randn(896, generator=g)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, 256, 2).float() / 256))
    ref_t = O.pos_encoding(t)
    ref_t = ref_t + torch.where((y >= 0)[:, None], label[y.clamp_min(0)], torch.zeros(1))
    ref_e = F.linear(F.silu(ref_t.double()), w.double(), b.double())
    temb = torch.empty(rows, 256, device=DEV)
    emb = torch.empty(rows, 896, device=DEV)
    ops.time_embed(t.to(DEV), None, y.to(DEV), inv_freq.to(DEV), label.to(DEV), w.to(DEV), b.to(DEV), temb, emb)
    # sin/cos of arguments up to 999 rad: device sinf vs host sinf differ by <= 2 ulp of a value <= 1
    assert float((temb.cpu() - ref_t).abs().max()) < 5e-7
    assert O.rel_l2(emb.cpu(), ref_e) < 2e-6
    # device step counter path == t path
    step = torch.tensor([500], dtype=torch.int32, device=DEV)
    temb2 = torch.empty(rows, 256, device=DEV)
    ops.time_embed(None, step, None, inv_freq.to(DEV), None, w.to(DEV), b.to(DEV), temb2, emb)
    assert float((temb2.cpu() - O.pos_encoding(torch.full((rows,), 500.0))).abs().max()) < 5e-7


def test_row_broadcast_of_the_shared_prefix(ops):
    """CFG batching computes the label-independent prefix (inc, down1's convs) once for n rows: GroupNorm-apply and
    upsample+concat then read raw / skip row r % n for output row r.  Must equal the same call on repeated inputs."""
    g = gen(31)
    n, H, C = 3, 8, 64
    raw = torch.randn(n, H, H, C, generator=g).to(DEV)
    part = torch.stack([raw.double().sum((1, 2, 3)), (raw.double() ** 2).sum((1, 2, 3))], -1).float().reshape(n, 1, 2).contiguous()
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    emb = torch.randn(2 * n, C, generator=g).to(DEV)
    for raw_t in (raw, raw.half()):
        out_b = torch.empty(2 * n, H, H, C, device=DEV)
        ops.gn_apply(raw_t, part, gamma, beta, mode=0, emb=emb, out_f32=out_b)
        out_r = torch.empty(2 * n, H, H, C, device=DEV)
        ops.gn_apply(raw_t.repeat(2, 1, 1, 1), part.repeat(2, 1, 1), gamma, beta, mode=0, emb=emb, out_f32=out_r)
        assert torch.equal(out_b, out_r)
        assert not torch.equal(out_b[:n], out_b[n:])  # the two halves differ by their embeddings
    x = torch.randn(2 * n, 4, 4, 64, generator=g).to(DEV)
    skip = torch.randn(n, 8, 8, 64, generator=g).to(DEV)
    o_b = torch.empty(2 * n, 8, 8, 128, device=DEV)
    o_r = torch.empty(2 * n, 8, 8, 128, device=DEV)
    ops.upsample_cat(x, skip, out_f32=o_b)
    ops.upsample_cat(x, skip.repeat(2, 1, 1, 1), out_f32=o_r)
    assert torch.equal(o_b, o_r)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,skip_rows,h,cx,cs", [(2, 2, 4, 64, 64), (4, 2, 8, 128, 128), (2, 1, 16, 64, 64), (3, 3, 2, 256, 256), (1, 1, 1, 64, 64)])
def test_gn_apply_vcat(ops, rows, skip_rows, h, cx, cs, dtype):
    """GELU(GroupNorm(raw) + cat([skip, upsample(x)])) with the residual recomputed from its sources equals
    sg_upsample_cat (fp32) followed by sg_gn_apply mode 2, bit for bit, and fp64 torch within 16-bit rounding."""
    g = gen(41)
    C = cx + cs
    x = torch.randn(rows, h, h, cx, generator=g).to(DEV)
    skip = torch.randn(skip_rows, 2 * h, 2 * h, cs, generator=g).to(DEV)
    raw = torch.randn(rows, 2 * h, 2 * h, C, generator=g).half().to(DEV)
    part = torch.stack([raw.double().sum((1, 2, 3)), (raw.double() ** 2).sum((1, 2, 3))], -1).float().reshape(rows, 1, 2).contiguous()
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    got = torch.empty(rows, 2 * h, 2 * h, C, device=DEV, dtype=dtype)
    ops.gn_apply_vcat(raw, part, gamma, beta, x, skip, got)
    cat = torch.empty(rows, 2 * h, 2 * h, C, device=DEV)
    ops.upsample_cat(x, skip, out_f32=cat)
    two = torch.empty_like(got)
    ops.gn_apply(raw, part, gamma, beta, mode=2, residual=cat, out_act=two)
    torch.cuda.synchronize()
    assert torch.equal(got, two)
    up = F.interpolate(nchw(x.cpu()).double(), scale_factor=2, mode="bilinear", align_corners=True)
    res = nhwc(torch.cat([nchw(skip.cpu().repeat(rows // skip_rows, 1, 1, 1)).double(), up], 1))
    gn = nhwc(F.group_norm(nchw(raw.cpu().double()), 1, gamma.cpu().double(), beta.cpu().double(), 1e-5))
    assert O.rel_l2(got.cpu(), F.gelu(gn + res)) < (4e-3 if dtype == torch.bfloat16 else 6e-4)


# ------------------------------------------------------------------------------------------ convolutions
@pytest.mark.parametrize("c_in,S,rows,n_src", [(4, 16, 4, 2), (1, 32, 2, 2), (4, 64, 2, 1), (3, 16, 1, 1)])
def test_conv_in(ops, c_in, S, rows, n_src):
    g = gen(5)
    x = torch.randn(n_src, c_in, S, S, generator=g)
    w = torch.randn(64, c_in, 3, 3, generator=g) / 3
    ref = nhwc(F.conv2d(x.double(), w.double(), padding=1))
    ref = ref[torch.arange(rows) % n_src]
    raw = torch.empty(rows, S, S, 64, device=DEV)
    part = torch.empty(rows, ops.conv_in_partials(S), 2, device=DEV)
    ops.conv_in(x.to(DEV), w, raw, part)  # w: host tensor (passed by value as a launch parameter)
    assert O.rel_l2(raw.cpu(), ref) < 1e-6
    s = part.cpu().double().sum(1)
    assert torch.allclose(s[:, 0], ref.sum((1, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[:, 1], (ref ** 2).sum((1, 2, 3)), rtol=1e-5)
    # fp16 raw output (tensor-core modes): same fp32 statistics, stored value rounded to 11 bits
    raw16 = torch.empty(rows, S, S, 64, device=DEV, dtype=torch.float16)
    part16 = torch.empty_like(part)
    ops.conv_in(x.to(DEV), w, raw16, part16)
    assert O.rel_l2(raw16.cpu(), ref) < 4e-4
    assert torch.equal(part16, part)


CONV_CASES = [  # rows, H, Cin, Cout
    (2, 16, 64, 64), (4, 2, 256, 256), (3, 4, 256, 512), (2, 8, 128, 128), (1, 32, 128, 64), (2, 16, 192, 128),
    (40, 2, 64, 64),
]


def _conv_ref(a, w):
    return nhwc(F.conv2d(nchw(a.double()), w.double(), padding=1))


@pytest.mark.parametrize("rows,H,cin,cout", CONV_CASES)
def test_igemm_conv_simt_with_groupnorm(ops, rows, H, cin, cout):
    from spectrogramgenai_b200._cabi import SG_ENGINE_SIMT

    g = gen(6)
    a = torch.randn(rows, H, H, cin, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    ref = _conv_ref(a, w)
    raw = torch.empty(rows, H, H, cout, device=DEV)
    part = torch.full((rows, ops.igemm_partials(SG_ENGINE_SIMT, H, H, cout), 2), float("nan"), device=DEV)
    ops.igemm(a.to(DEV), pack_conv(w).to(DEV), rows=rows, H=H, W=H, out_f32=raw, partials=part)
    assert O.rel_l2(raw.cpu(), ref) < 2e-6
    # GroupNorm(1, C) finalize + affine + GELU / residual / embedding
    gamma = torch.randn(cout, generator=g)
    beta = torch.randn(cout, generator=g)
    res = torch.randn(rows, H, H, cout, generator=g)
    emb = torch.randn(rows, cout + 64, generator=g)
    gn = nhwc(F.group_norm(nchw(ref), 1, gamma.double(), beta.double(), 1e-5))
    for mode, want in ((0, gn), (1, F.gelu(gn)), (2, F.gelu(gn + res.double()))):
        out = torch.empty(rows, H, H, cout, device=DEV)
        ops.gn_apply(raw, part, gamma.to(DEV), beta.to(DEV), mode=mode, residual=res.to(DEV) if mode == 2 else None,
                     out_f32=out)
        assert O.rel_l2(out.cpu(), want) < 3e-6, mode
    out = torch.empty(rows, H, H, cout, device=DEV)
    o16 = torch.empty(rows, H, H, cout, device=DEV, dtype=torch.bfloat16)
    embd = emb.to(DEV)
    ops.gn_apply(raw, part, gamma.to(DEV), beta.to(DEV), mode=0, emb=embd[:, 64:], out_f32=out, out_act=o16)
    want = gn + emb[:, 64:].double()[:, None, None, :]
    assert O.rel_l2(out.cpu(), want) < 3e-6
    assert O.rel_l2(o16.cpu(), want) < 4e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
# (.., 64) at H = 64 / 32: the slab pipeline of the Cout = 64 layers (one TMA box per column shift serves three taps); the
# 16-row case has more 256-pixel tiles (256) than SMs, the 3- and 5-row cases an uneven last wave
@pytest.mark.parametrize("rows,H,cin,cout", CONV_CASES + [(2, 64, 64, 64), (1, 64, 128, 128), (3, 64, 128, 64), (5, 32, 64, 64),
                                                          (16, 64, 64, 64), (2, 32, 192, 64)])
def test_igemm_conv_tensor_core(ops, rows, H, cin, cout, dtype):
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    g = gen(7)
    a = torch.randn(rows, H, H, cin, generator=g).to(dtype)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(dtype)
    ref = _conv_ref(a.float(), w.float())  # fp64 on the rounded operands
    raw = torch.full((rows, H, H, cout), float("nan"), device=DEV)
    o16 = torch.empty(rows, H, H, cout, device=DEV, dtype=dtype)
    part = torch.full((rows, ops.igemm_partials(SG_ENGINE_TC, H, H, cout), 2), float("nan"), device=DEV)
    ops.igemm(a.to(DEV), pack_conv(w.float(), dtype).to(DEV), rows=rows, H=H, W=H, out_f32=raw, out_act=o16,
              partials=part)
    torch.cuda.synchronize()
    assert O.rel_l2(raw.cpu(), ref) < 5e-6  # fp32 accumulation over K up to 9*256
    assert O.rel_l2(o16.cpu(), ref) < 4e-3
    s = part.cpu().double().sum(1)
    assert torch.allclose(s[:, 0], ref.sum((1, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[:, 1], (ref ** 2).sum((1, 2, 3)), rtol=1e-5)



@pytest.mark.parametrize("scale,expect", [(1.0, 0), (3000.0, 1), (2e-4, 1), (30.0, 0), (0.0, 0)])
def test_gn_apply_fp16_range_flag(ops, scale, expect):
    """GroupNorm is scale invariant, its fp16 raw input is not: the kernel compares the exact mean square (fp32 statistics)
    with [2^-20, 2^20] and raises the flag the host uses to fall back to fp32 raw tensors.  All-zero rows are fine."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    rows, H, cin, cout = 2, 16, 64, 64
    g = gen(45)
    a = torch.randn(rows, H, H, cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin) * scale).to(torch.bfloat16)
    raw = torch.empty(rows, H, H, cout, device=DEV, dtype=torch.float16)
    part = torch.empty(rows, ops.igemm_partials(SG_ENGINE_TC, H, H, cout), 2, device=DEV)
    ops.igemm(a.to(DEV), pack_conv(w.float(), torch.bfloat16).to(DEV), rows=rows, H=H, W=H, out_act=raw, partials=part)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = torch.empty(rows, H, H, cout, device=DEV, dtype=torch.bfloat16)
    ones, zeros = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    ops.gn_apply(raw, part, ones, zeros, mode=1, out_act=out, range_flag=flag)
    assert int(flag.item()) == expect

def test_gn_apply_fp16_raw_gelu_accuracy(ops):
    """The 16-bit engines' GroupNorm-apply evaluates GELU as x / (1 + 2^(x p(x^2))) (gelu_logistic2, 8 packed instructions
    per pair): against the exact erf GELU of the same normalised values the fp32 output is within 5e-6 absolute over
    |x| <= 13, on both fp16-raw code paths (fixed-channel and generic)."""
    for C, HW in ((64, 4096), (24, 1000)):  # 24 channels: the generic path (grid stride not a multiple of C / 8)
        raw = torch.linspace(-3.0, 3.0, HW * C).reshape(1, HW, C).half()
        x = raw.double()
        part = torch.stack([x.sum((1, 2)), (x * x).sum((1, 2))], -1).float().reshape(1, 1, 2).contiguous().to(DEV)
        gamma, beta = torch.full((C,), 7.5), torch.zeros(C)
        norm = (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-5) * 7.5
        out = torch.empty(1, HW, C, device=DEV)
        ops.gn_apply(raw.to(DEV), part, gamma.to(DEV), beta.to(DEV), mode=1, out_f32=out)
        err = (out.cpu().double() - F.gelu(norm)).abs().max().item()
        print(f"gn_apply fp16-raw GELU C={C}: max |error| {err:.3e} over x in [{norm.min():.1f}, {norm.max():.1f}]")
        assert norm.max() > 12 and err < 5e-6 + 2e-6  # + the rounding of the fp32 normalisation itself at |x| ~ 13


# ---------------------------------------------------------------- fp32-accurate tensor-core engine (split TF32)
def test_split_tf32(ops):
    """hi = tf32(x) (10 explicit mantissa bits), lo = tf32(x - hi); hi + lo reproduces x to ~2^-22."""
    x = torch.randn(4096 * 4, generator=gen(40)) * torch.logspace(-6, 6, 4096 * 4)
    hi, lo = ops.split_tf32(x.to(DEV))
    hi, lo = hi.cpu(), lo.cpu()
    assert (hi.view(torch.int32) & 0x1FFF).eq(0).all() and (lo.view(torch.int32) & 0x1FFF).eq(0).all()
    assert ((hi - x).abs() <= x.abs() * 2.0 ** -11 + 1e-45).all()
    assert (((hi.double() + lo.double()) - x.double()).abs() <= x.abs().double() * 2.0 ** -21).all()


def _trunc19(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def test_split_tf32_lo_and_producers(ops):
    """Activation operand form: the fp32 tensor is its own high part (kind::tf32 reads its top 19 bits) and
    lo = tf32(x - trunc19(x)).  sg_split_tf32(hi = NULL) and the producers that write lo next to their fp32 output
    (GroupNorm-apply on fp32 raw, maxpool, upsample + concat) agree bit for bit."""
    g = gen(45)
    x = torch.randn(4096 * 4, generator=g) * torch.logspace(-6, 6, 4096 * 4)
    lo = ops.split_tf32_lo(x.to(DEV)).cpu()
    assert (lo.view(torch.int32) & 0x1FFF).eq(0).all()
    assert (((_trunc19(x).double() + lo.double()) - x.double()).abs() <= x.abs().double() * 2.0 ** -21).all()
    # GroupNorm-apply (fp32 raw), all three modes
    rows, H, C = 3, 8, 64
    raw = torch.randn(rows, H, H, C, generator=g).to(DEV)
    part = torch.stack([raw.sum((1, 2, 3)), (raw * raw).sum((1, 2, 3))], -1).reshape(rows, 1, 2).contiguous()
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    res = torch.randn(rows, H, H, C, generator=g).to(DEV)
    for mode in (0, 1, 2):
        o32, ol = torch.empty_like(raw), torch.empty_like(raw)
        ops.gn_apply(raw, part, gamma, beta, mode=mode, residual=res if mode == 2 else None, out_f32=o32, out_act=ol)
        assert torch.equal(ol, ops.split_tf32_lo(o32))
        only = torch.empty_like(raw)
        ops.gn_apply(raw, part, gamma, beta, mode=mode, residual=res if mode == 2 else None, out_f32=only)
        assert torch.equal(only, o32)
    x4 = torch.randn(2, 16, 16, 64, generator=g).to(DEV)
    p32, pl = torch.empty(2, 8, 8, 64, device=DEV), torch.empty(2, 8, 8, 64, device=DEV)
    ops.maxpool2(x4, out_f32=p32, out_act=pl)
    assert torch.equal(pl, ops.split_tf32_lo(p32))
    for cx, cs in ((64, 64), (32, 64), (4, 12)):  # paired / 8-wide / generic kernels
        xs, sk = torch.randn(2, 8, 8, cx, generator=g).to(DEV), torch.randn(2, 16, 16, cs, generator=g).to(DEV)
        c32, cl = torch.empty(2, 16, 16, cx + cs, device=DEV), torch.empty(2, 16, 16, cx + cs, device=DEV)
        ops.upsample_cat(xs, sk, out_f32=c32, out_act=cl)
        assert torch.equal(cl, ops.split_tf32_lo(c32))


@pytest.mark.parametrize("rows,H,cin,cout", [(2, 16, 64, 128), (1, 64, 128, 128), (2, 8, 512, 512)])
def test_igemm_conv_split_tf32_activation_form(ops, rows, H, cin, cout):
    """The same conv with the activation given as (x, tf32_lo(x)) -- the form the engine uses -- is as accurate as with the
    rounded (hi, lo) pair."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    g = gen(46)
    a = torch.randn(rows, H, H, cin, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    ref = _conv_ref(a, w)
    ad = a.to(DEV)
    wp = ops.split_tf32(pack_conv(w).to(DEV))
    outs = []
    for a_op in ((ad, ops.split_tf32_lo(ad)), ops.split_tf32(ad)):
        raw = torch.full((rows, H, H, cout), float("nan"), device=DEV)
        part = torch.full((rows, ops.igemm_partials(SG_ENGINE_TC, H, H, cout), 2), float("nan"), device=DEV)
        ops.igemm(a_op, wp, rows=rows, H=H, W=H, out_f32=raw, partials=part)
        got = raw.cpu().double()
        scale = float((got * ref).sum() / (ref * ref).sum())
        outs.append(O.rel_l2(got / scale, ref))
    print(f"igemm split-tf32 {cin}->{cout} H={H}: activation form {outs[0]:.3e}, rounded pair {outs[1]:.3e}")
    assert outs[0] < 8e-6 and outs[0] < 1.2 * outs[1] + 5e-7  # (the K = 4608 case sits at 5e-6 in both forms: accumulation)


@pytest.mark.parametrize("rows,H,cin,cout", CONV_CASES + [(2, 64, 64, 64), (1, 64, 128, 128), (2, 16, 96, 64)])
def test_igemm_conv_split_tf32(ops, rows, H, cin, cout):
    """3x3 conv on the split-TF32 engine (three kind::tf32 MMAs per product) vs fp64 on the SAME fp32 operands: the bar is
    the fp32 engine's (2e-6 here; a single TF32 pass gives ~5e-4), and the GroupNorm partials come with it."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    g = gen(41)
    a = torch.randn(rows, H, H, cin, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    ref = _conv_ref(a, w)
    raw = torch.full((rows, H, H, cout), float("nan"), device=DEV)
    part = torch.full((rows, ops.igemm_partials(SG_ENGINE_TC, H, H, cout), 2), float("nan"), device=DEV)
    ops.igemm(ops.split_tf32(a.to(DEV)), ops.split_tf32(pack_conv(w).to(DEV)), rows=rows, H=H, W=H, out_f32=raw,
              partials=part)
    torch.cuda.synchronize()
    got = raw.cpu().double()
    err = O.rel_l2(got, ref)
    # The tensor core accumulates in fp32 with truncation: a uniform relative shrink of ~2.5e-8 per accumulating MMA of the
    # hi x hi pass (K / 8 of them), i.e. a scale factor 1 - O(1e-8 K) on the whole output -- which the GroupNorm behind
    # every 3x3 conv of the model removes.  Apart from that factor the result is fp32-accurate.
    scale = float((got * ref).sum() / (ref * ref).sum())
    err_scaled = O.rel_l2(got / scale, ref)
    print(f"igemm split-tf32 rows={rows} H={H} {cin}->{cout}: rel-L2 {err:.3e}, scale 1{scale - 1:+.2e}, after scale {err_scaled:.3e}")
    assert err < 2e-6 + 1.5e-8 * 9 * cin and abs(scale - 1) < 1.5e-8 * 9 * cin + 1e-6
    assert err_scaled < 4e-6
    s = part.cpu().double().sum(1)
    assert torch.allclose(s[:, 0], ref.sum((1, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[:, 1], (ref ** 2).sum((1, 2, 3)), rtol=1e-5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,H,cin,cout", [(2, 16, 64, 128), (3, 8, 256, 256), (1, 64, 64, 64)])
def test_conv_fp16_raw_then_groupnorm(ops, rows, H, cin, cout, dtype):
    """The tensor-core modes keep the raw conv output in fp16 whatever the operand type (sg_igemm out_dtype): the
    GroupNorm statistics still come from the fp32 accumulators, the stored value carries 2^-12 relative rounding,
    and a value beyond the fp16 range saturates instead of becoming inf."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    g = gen(17)
    a = torch.randn(rows, H, H, cin, generator=g).to(dtype)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(dtype)
    ref = _conv_ref(a.float(), w.float())
    raw16 = torch.full((rows, H, H, cout), float("nan"), device=DEV, dtype=torch.float16)
    part = torch.full((rows, ops.igemm_partials(SG_ENGINE_TC, H, H, cout), 2), float("nan"), device=DEV)
    ops.igemm(a.to(DEV), pack_conv(w.float(), dtype).to(DEV), rows=rows, H=H, W=H, out_act=raw16, partials=part)
    torch.cuda.synchronize()
    assert O.rel_l2(raw16.cpu(), ref) < 4e-4
    s = part.cpu().double().sum(1)
    assert torch.allclose(s[:, 1], (ref ** 2).sum((1, 2, 3)), rtol=1e-5)  # statistics are those of the fp32 result
    gamma, beta = torch.randn(cout, generator=g), torch.randn(cout, generator=g)
    res = torch.randn(rows, H, H, cout, generator=g)
    gn = nhwc(F.group_norm(nchw(ref), 1, gamma.double(), beta.double(), 1e-5))
    for mode, want in ((0, gn), (1, F.gelu(gn)), (2, F.gelu(gn + res.double()))):
        out = torch.empty(rows, H, H, cout, device=DEV)
        ops.gn_apply(raw16, part, gamma.to(DEV), beta.to(DEV), mode=mode, residual=res.to(DEV) if mode == 2 else None,
                     out_f32=out)
        assert O.rel_l2(out.cpu(), want) < 6e-4, mode
    # saturation: weights scaled so that the result leaves the fp16 range
    big = torch.full((1, 8, 8, 64), 200.0).to(dtype)
    wbig = torch.full((64, 64, 3, 3), 1.0).to(dtype)
    rawb = torch.empty(1, 8, 8, 64, device=DEV, dtype=torch.float16)
    partb = torch.empty(1, ops.igemm_partials(SG_ENGINE_TC, 8, 8, 64), 2, device=DEV)
    ops.igemm(big.to(DEV), pack_conv(wbig.float(), dtype).to(DEV), rows=1, H=8, W=8, out_act=rawb, partials=partb)
    torch.cuda.synchronize()
    assert torch.isfinite(rawb.float()).all() and float(rawb.float().max()) == 65504.0


LIN_CASES = [(2, 4, 64, 192), (2, 8, 128, 384), (1, 16, 256, 768), (3, 2, 256, 256), (2, 32, 64, 64)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,H,cin,cout", LIN_CASES)
def test_igemm_linear_epilogues(ops, rows, H, cin, cout, dtype):
    g = gen(8)
    M = rows * H * H
    a = torch.randn(M, cin, generator=g).to(dtype)
    w = (torch.randn(cout, cin, generator=g) / math.sqrt(cin)).to(dtype)
    b = torch.randn(cout, generator=g)
    res = torch.randn(M, cout, generator=g)
    lin = F.linear(a.double(), w.double(), b.double())
    wp = w.reshape(1, cout, cin).contiguous().to(DEV)
    out = torch.empty(M, cout, device=DEV)
    ops.igemm(a.to(DEV), wp, rows=rows, H=H, W=H, bias=b.to(DEV), residual=res.to(DEV), out_f32=out)
    assert O.rel_l2(out.cpu(), lin + res.double()) < 3e-6
    if dtype == torch.float32:
        ops.igemm(a.to(DEV), wp, rows=rows, H=H, W=H, bias=b.to(DEV), gelu=True, out_f32=out)
        assert O.rel_l2(out.cpu(), F.gelu(lin)) < 3e-6
    else:
        o16 = torch.empty(M, cout, device=DEV, dtype=dtype)
        ops.igemm(a.to(DEV), wp, rows=rows, H=H, W=H, bias=b.to(DEV), gelu=True, out_act=o16)
        assert O.rel_l2(o16.cpu(), F.gelu(lin)) < 4e-3


@pytest.mark.parametrize("rows,H,cin,cout", LIN_CASES)
def test_igemm_linear_split_tf32(ops, rows, H, cin, cout):
    g = gen(42)
    M = rows * H * H
    a = torch.randn(M, cin, generator=g)
    w = torch.randn(cout, cin, generator=g) / math.sqrt(cin)
    b = torch.randn(cout, generator=g)
    res = torch.randn(M, cout, generator=g)
    lin = F.linear(a.double(), w.double(), b.double())
    wp = ops.split_tf32(w.reshape(1, cout, cin).contiguous().to(DEV))
    ap = ops.split_tf32(a.to(DEV))
    out = torch.empty(M, cout, device=DEV)
    ops.igemm(ap, wp, rows=rows, H=H, W=H, bias=b.to(DEV), residual=res.to(DEV), out_f32=out)
    assert O.rel_l2(out.cpu(), lin + res.double()) < 4e-6
    ops.igemm(ap, wp, rows=rows, H=H, W=H, bias=b.to(DEV), gelu=True, out_f32=out)
    assert O.rel_l2(out.cpu(), F.gelu(lin)) < 5e-6


def test_conv_out(ops):
    g = gen(9)
    for c_out, rows, S in ((4, 3, 16), (1, 2, 32)):
        x = torch.randn(rows, S * S, 64, generator=g)
        w = torch.randn(c_out, 64, generator=g) / 8
        b = torch.randn(c_out, generator=g)
        ref = (x.double() @ w.double().T + b.double()).transpose(1, 2).reshape(rows, c_out, S, S)
        eps = torch.empty(rows, c_out, S, S, device=DEV)
        ops.conv_out(x.to(DEV), w.to(DEV), b.to(DEV), eps)
        assert O.rel_l2(eps.cpu(), ref) < 1e-6


# ------------------------------------------------------------------------------------------ norms / attention
@pytest.mark.parametrize("M", [37, 1, 64, 4099])
@pytest.mark.parametrize("C", [64, 128, 256])
def test_layernorm(ops, C, M):
    """Ragged token counts: every lane group / unrolled pass of the vectorised kernel sees a partial tail."""
    g = gen(10)
    x = torch.randn(M, C, generator=g) * 3 + 1
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.layer_norm(x.double(), (C,), gamma.double(), beta.double(), 1e-5)
    for dt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 4e-3), (torch.float16, 6e-4)):
        out = torch.full((M + 3, C), float("nan"), device=DEV, dtype=dt)
        ops.layernorm(x.to(DEV), gamma.to(DEV), beta.to(DEV), out[:M])
        assert O.rel_l2(out[:M].cpu(), ref) < tol
        assert torch.isnan(out[M:].float()).all()  # nothing written past the last token


def _sa_weights(C, g):
    w = lambda *s: torch.randn(*s, generator=g) / s[-1] ** 0.5  # noqa: E731
    return dict(ln_g=1 + 0.2 * torch.randn(C, generator=g), ln_b=0.2 * torch.randn(C, generator=g),
                w_in=w(3 * C, C), b_in=0.3 * torch.randn(3 * C, generator=g),
                wo=w(C, C), bo=0.3 * torch.randn(C, generator=g),
                ln2_g=1 + 0.2 * torch.randn(C, generator=g), ln2_b=0.2 * torch.randn(C, generator=g),
                w1=w(C, C), b1=0.3 * torch.randn(C, generator=g), w2=w(C, C), b2=0.3 * torch.randn(C, generator=g))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C", [64, 128])
@pytest.mark.parametrize("M", [128, 4096 + 128, 200, 100000])
def test_ln_inproj_fused(ops, M, C, dtype):
    """Fused LayerNorm + in_proj vs fp64 torch on the same 16-bit weights.  Rounding added by the kernel: LN output
    and qkv to 16 bit (what the unfused path does too).  M = 200 exercises the TMA-clipped last tile."""
    g = gen(21)
    W = _sa_weights(C, g)
    x = torch.randn(M, C, generator=g) * 2 + 0.5
    w16 = W["w_in"].to(dtype)
    ln = F.layer_norm(x.double(), (C,), W["ln_g"].double(), W["ln_b"].double(), 1e-5)
    ref = ln @ w16.double().T + W["b_in"].double()
    qkv = torch.full((M + 1, 3 * C), float("nan"), device=DEV, dtype=dtype)
    ops.ln_inproj(x.to(DEV), W["ln_g"].to(DEV), W["ln_b"].to(DEV), w16.to(DEV), W["b_in"].to(DEV), qkv[:M])
    torch.cuda.synchronize()
    err = O.rel_l2(qkv[:M].cpu(), ref)
    print(f"ln_inproj M={M} C={C} {dtype}: rel-L2 {err:.3e}")
    assert err < (6e-3 if dtype == torch.bfloat16 else 8e-4)
    assert torch.isnan(qkv[M:].float()).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C", [64, 128])
@pytest.mark.parametrize("M", [128, 4096 + 128, 200, 100000])
def test_attn_tail_fused(ops, M, C, dtype):
    """Fused out_proj + residual + LayerNorm + FFN + residual vs fp64 torch on the same 16-bit att / weights."""
    g = gen(22)
    W = _sa_weights(C, g)
    x = torch.randn(M, C, generator=g) * 2
    att = torch.randn(M, C, generator=g).to(dtype)
    wo, w1, w2 = (W[k].to(dtype) for k in ("wo", "w1", "w2"))
    a = att.double() @ wo.double().T + W["bo"].double() + x.double()
    h = F.gelu(F.layer_norm(a, (C,), W["ln2_g"].double(), W["ln2_b"].double(), 1e-5) @ w1.double().T + W["b1"].double())
    ref = h @ w2.double().T + W["b2"].double() + a
    out = torch.full((M + 1, C), float("nan"), device=DEV)
    d = lambda t: t.to(DEV)  # noqa: E731
    ops.attn_tail(d(att), d(x), d(wo), d(W["bo"]), d(W["ln2_g"]), d(W["ln2_b"]), d(w1), d(W["b1"]), d(w2), d(W["b2"]),
                  out[:M])
    torch.cuda.synchronize()
    err = O.rel_l2(out[:M].cpu(), ref)
    print(f"attn_tail M={M} C={C} {dtype}: rel-L2 {err:.3e}")
    assert err < (3e-3 if dtype == torch.bfloat16 else 4e-4)
    assert torch.isnan(out[M:]).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,HW,c_out", [(3, 256, 4), (2, 1024, 1), (5, 128, 3), (1100, 256, 4)])
def test_attn_tail_with_fused_output_conv(ops, rows, HW, c_out, dtype):
    """sg_attn_tail_outc: eps NCHW = outc(block output), computed from the fp32 rows the tail holds.  The block output
    it optionally also writes is bit-identical to sg_attn_tail's, eps is the same with and without that write, and
    eps matches the 1x1 conv of that output (fp64) to fp32 rounding."""
    C, M = 64, rows * HW
    g = gen(23)
    W = _sa_weights(C, g)
    x = torch.randn(M, C, generator=g) * 2
    att = torch.randn(M, C, generator=g).to(dtype)
    wc = torch.randn(c_out, C, generator=g) * 0.2
    bc = torch.randn(c_out, generator=g)
    d = lambda t: t.to(DEV)  # noqa: E731
    args = [d(att), d(x)] + [d(t) for t in (W["wo"].to(dtype), W["bo"], W["ln2_g"], W["ln2_b"], W["w1"].to(dtype), W["b1"],
                                            W["w2"].to(dtype), W["b2"])]
    plain = torch.empty(M, C, device=DEV)
    ops.attn_tail(*args, plain)
    both = torch.empty(M, C, device=DEV)
    eps1 = torch.full((rows + 1, c_out, HW // 16, 16), float("nan"), device=DEV)
    ops.attn_tail(*args, both, outc=(wc, bc, eps1[:rows]))  # outc weights: host tensors (launch parameters)
    eps2 = torch.empty((rows, c_out, HW // 16, 16), device=DEV)
    ops.attn_tail(*args, None, outc=(wc, bc, eps2))
    torch.cuda.synchronize()
    assert torch.equal(plain, both)
    assert torch.equal(eps1[:rows], eps2) and torch.isnan(eps1[rows:]).all()
    ref = (plain.cpu().double() @ wc.double().T + bc.double()).reshape(rows, HW, c_out).permute(0, 2, 1)
    err = O.rel_l2(eps2.cpu().reshape(rows, c_out, HW), ref)
    print(f"attn_tail+outc rows={rows} HW={HW} c_out={c_out} {dtype}: rel-L2 {err:.3e}")
    assert err < 2e-6
    with pytest.raises(Exception):  # C = 128 has no fused output conv
        ops.attn_tail(d(att.reshape(-1, 128)), d(x.reshape(-1, 128)), *args[2:], None,
                      outc=(wc, bc, eps2))


def _attention_ref(qkv, rows, L, C):
    d = C // 4
    q, k, v = qkv.double().reshape(rows, L, 3 * C).split(C, -1)
    h = lambda z: z.reshape(rows, L, 4, d).transpose(1, 2)  # noqa: E731
    att = torch.softmax(h(q) * d ** -0.5 @ h(k).transpose(-1, -2), -1) @ h(v)
    return att.transpose(1, 2).reshape(rows * L, C)


@pytest.mark.parametrize("rows,L,C", [(2, 4, 256), (3, 16, 256), (2, 64, 256), (2, 256, 128), (1, 1024, 64), (2, 200, 64)])
def test_attention_simt(ops, rows, L, C):
    from spectrogramgenai_b200._cabi import SG_ENGINE_SIMT

    qkv = torch.randn(rows * L, 3 * C, generator=gen(11)) * 1.5
    ref = _attention_ref(qkv, rows, L, C)
    out = torch.empty(rows * L, C, device=DEV)
    ops.attention(qkv.to(DEV), out, rows=rows, L=L, C=C, engine=SG_ENGINE_SIMT)
    assert O.rel_l2(out.cpu(), ref) < 3e-6
    o16 = torch.empty(rows * L, C, device=DEV, dtype=torch.float16)
    ops.attention(qkv.to(DEV), o16, rows=rows, L=L, C=C, engine=SG_ENGINE_SIMT)
    assert O.rel_l2(o16.cpu(), ref) < 6e-4


ATT_TC_CASES = [(2, 4, 256), (5, 16, 256), (3, 64, 256), (2, 256, 256), (2, 256, 128), (2, 1024, 128), (1, 1024, 64),
                (2, 4096, 64), (40, 4, 256), (1, 16, 128)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,L,C", ATT_TC_CASES)
def test_attention_tensor_core(ops, rows, L, C, dtype):
    """tcgen05 flash attention vs fp64 softmax(QK^T/sqrt(d))V on the same 16-bit q/k/v.  The only rounding the
    kernel adds is P -> 16 bit before P V (2^-9 bf16 / 2^-12 f16 per element) and the 16-bit output."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    qkv = (torch.randn(rows * L, 3 * C, generator=gen(12)) * 1.5).to(dtype)
    ref = _attention_ref(qkv.float(), rows, L, C)
    out = torch.full((rows * L, C), float("nan"), device=DEV, dtype=dtype)
    ops.attention(qkv.to(DEV), out, rows=rows, L=L, C=C, engine=SG_ENGINE_TC)
    torch.cuda.synchronize()
    err = O.rel_l2(out.cpu(), ref)
    print(f"attention tc rows={rows} L={L} C={C} {dtype}: rel-L2 {err:.3e}")
    assert err < (6e-3 if dtype == torch.bfloat16 else 1e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,L,C", [(2, 1024, 64), (1, 512, 128), (1, 256, 256)])
def test_attention_tensor_core_growing_scores(ops, rows, L, C, dtype):
    """Keys whose magnitude grows along the sequence: the running maximum rises by far more than 2^8 between key
    tiles, which exercises the lazy-rescale path (O and l rescaled in TMEM) of the two-tile kernel."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    qkv = torch.randn(rows * L, 3 * C, generator=gen(13))
    ramp = (1.0 + 24.0 * torch.arange(L) / L).repeat(rows)[:, None]
    qkv[:, C:2 * C] *= ramp  # K
    qkv = qkv.to(dtype)
    ref = _attention_ref(qkv.float(), rows, L, C)
    out = torch.full((rows * L, C), float("nan"), device=DEV, dtype=dtype)
    ops.attention(qkv.to(DEV), out, rows=rows, L=L, C=C, engine=SG_ENGINE_TC)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    err = O.rel_l2(out.cpu(), ref)
    print(f"attention tc (growing scores) rows={rows} L={L} C={C} {dtype}: rel-L2 {err:.3e}")
    assert err < (8e-3 if dtype == torch.bfloat16 else 1.5e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,L,C", [(2, 512, 64), (1, 4096, 64), (1, 256, 128)])
def test_attention_tensor_core_reference_overflow(ops, rows, L, C, dtype):
    """Scores whose spread exceeds the dynamic range of the fast pass (the d = 16 kernel takes the maximum of a row's
    first 16 scores as the softmax reference and never searches for it again): later keys beat that reference by far
    more than 2^60, the tile row sums overflow, and the query tile must be recomputed by the safe pass (full-tile
    maximum before every sweep).  Rows are nearly one-hot, so the result is essentially one V row."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    qkv = torch.randn(rows * L, 3 * C, generator=gen(14))
    ramp = (0.05 + 60.0 * (torch.arange(L) / L) ** 2).repeat(rows)[:, None]  # tiny first keys, huge last ones
    qkv[:, C:2 * C] *= ramp
    qkv[:, :C] *= 3.0
    qkv = qkv.to(dtype)
    ref = _attention_ref(qkv.float(), rows, L, C)
    out = torch.full((rows * L, C), float("nan"), device=DEV, dtype=dtype)
    ops.attention(qkv.to(DEV), out, rows=rows, L=L, C=C, engine=SG_ENGINE_TC)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    err = O.rel_l2(out.cpu(), ref)
    print(f"attention tc (reference overflow) rows={rows} L={L} C={C} {dtype}: rel-L2 {err:.3e}")
    assert err < (8e-3 if dtype == torch.bfloat16 else 1.5e-3)


def test_attention_grid_split_beyond_32768_query_tiles(ops):
    """More than 32768 query tiles per head (n > 512 at [4, 64, 64]) spill into the grid's z dimension.  Property check at
    rows = 2048, L = 4096, d = 16 (65536 tiles): every row equals what it gets when its half is run alone."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    rows, L, C = 2048, 4096, 64
    g = torch.Generator(device=DEV).manual_seed(3)
    qkv = torch.empty(rows * L, 3 * C, device=DEV, dtype=torch.bfloat16)
    for r0 in range(0, rows, 256):  # generated in slices: no full-size fp32 temporary
        qkv[r0 * L:(r0 + 256) * L] = torch.randn(256 * L, 3 * C, device=DEV, generator=g).to(torch.bfloat16)
    out = torch.empty(rows * L, C, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, rows=rows, L=L, C=C, engine=SG_ENGINE_TC)
    half = rows // 2
    for h in (0, 1):
        o = torch.empty(half * L, C, device=DEV, dtype=torch.bfloat16)
        ops.attention(qkv[h * half * L:(h + 1) * half * L], o, rows=half, L=L, C=C, engine=SG_ENGINE_TC)
        assert torch.equal(o, out[h * half * L:(h + 1) * half * L]), h
    assert torch.isfinite(out.float()).all()



def _attention_tf32(ops, qkv, rows, L, C):
    M = rows * L
    qk_hi, qk_lo = torch.empty(M, 2 * C, device=DEV), torch.empty(M, 2 * C, device=DEV)
    vt_hi, vt_lo = torch.empty(rows * C, L, device=DEV), torch.empty(rows * C, L, device=DEV)
    ops.attn_prep_tf32(qkv.to(DEV), qk_hi, qk_lo, vt_hi, vt_lo, rows=rows, L=L, C=C)
    out = torch.full((M, C), float("nan"), device=DEV)
    ops.attention_tf32(qk_hi, qk_lo, vt_hi, vt_lo, out, rows=rows, L=L, C=C)
    torch.cuda.synchronize()
    return out.cpu(), (qk_hi.cpu(), qk_lo.cpu(), vt_hi.cpu(), vt_lo.cpu())


@pytest.mark.parametrize("rows,L,C", [(2, 256, 256), (2, 256, 128), (2, 1024, 128), (1, 1024, 64), (2, 4096, 64), (3, 4096, 64),
                                      (2, 128, 256), (5, 128, 64), (1, 128, 128)])
def test_attention_split_tf32(ops, rows, L, C):
    """The fp32-accurate attention core (S = Q K^T and O = P V as three kind::tf32 MMAs each on split operands, L >= 128)
    vs fp64 on the same fp32 q / k / v: the CUDA-core kernel's order of magnitude (1e-5; a single TF32 pass: ~1e-3).  The
    preparation pass is checked on its own: hi + lo reproduces q | k, and V comes out transposed per (row, head)."""
    qkv = torch.randn(rows * L, 3 * C, generator=gen(43)) * 1.5
    ref = _attention_ref(qkv, rows, L, C)
    out, (qk_hi, qk_lo, vt_hi, vt_lo) = _attention_tf32(ops, qkv, rows, L, C)
    assert ((qk_hi.double() + qk_lo.double()) - qkv[:, :2 * C].double()).abs().max() < 1e-6
    v = qkv[:, 2 * C:].reshape(rows, L, C).transpose(1, 2).reshape(rows * C, L)
    assert ((vt_hi.double() + vt_lo.double()) - v.double()).abs().max() < 1e-6
    assert (vt_hi.view(torch.int32) & 0x1FFF).eq(0).all() and (qk_lo.view(torch.int32) & 0x1FFF).eq(0).all()
    err = O.rel_l2(out, ref)
    print(f"attention split-tf32 rows={rows} L={L} C={C}: rel-L2 {err:.3e}")
    assert err < 1e-5


def test_attention_split_tf32_growing_scores(ops):
    """Scores that grow along the key axis force the lazily tracked softmax reference to be raised tile after tile."""
    rows, L, C = 2, 1024, 64
    g = gen(44)
    qkv = torch.randn(rows * L, 3 * C, generator=g)
    ramp = torch.linspace(0.2, 6.0, L).repeat(rows)[:, None]
    qkv[:, C:2 * C] *= ramp  # |k_j| grows with j
    ref = _attention_ref(qkv, rows, L, C)
    out, _ = _attention_tf32(ops, qkv, rows, L, C)
    assert O.rel_l2(out, ref) < 1e-5


def test_attention_ragged_grid_z(ops):
    """More than 32768 query tiles that are NOT a multiple of 32768 (e.g. 486 x 2 rows of L = 16384 in the generator at
    img_size 512): the last grid.z slice is ragged and its surplus CTAs exit.  rows = 1100, L = 4096 -> 35200 tiles."""
    from spectrogramgenai_b200._cabi import SG_ENGINE_TC

    rows, L, C = 1100, 4096, 64
    g = torch.Generator(device=DEV).manual_seed(4)
    qkv = torch.empty(rows * L, 3 * C, device=DEV, dtype=torch.bfloat16)
    for r0 in range(0, rows, 100):
        qkv[r0 * L:(r0 + 100) * L] = torch.randn(100 * L, 3 * C, device=DEV, generator=g).to(torch.bfloat16)
    out = torch.full((rows * L, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, rows=rows, L=L, C=C, engine=SG_ENGINE_TC)
    tail = torch.empty(76 * L, C, device=DEV, dtype=torch.bfloat16)  # rows 1024 .. 1099 live in the ragged slice
    ops.attention(qkv[1024 * L:], tail, rows=76, L=L, C=C, engine=SG_ENGINE_TC)
    assert torch.equal(tail, out[1024 * L:])
    assert torch.isfinite(out.float()).all()


def test_error_reporting(ops):
    """Bad arguments come back as a status + message (no exception crosses the C ABI, no crash)."""
    from spectrogramgenai_b200._cabi import SgError

    x = torch.zeros(2, 4, 24, 24, device=DEV)  # 24 is not a power of two
    with pytest.raises(SgError, match="power of two"):
        ops.conv_in(x, torch.zeros(64, 4, 3, 3), torch.zeros(2, 24, 24, 64, device=DEV),
                    torch.zeros(2, 9, 2, device=DEV))
    with pytest.raises(SgError, match="head dim"):
        ops.attention(torch.zeros(4, 3 * 32, device=DEV), torch.zeros(4, 32, device=DEV), rows=1, L=4, C=32)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(64, 4, 3, 3), (128, 64, 3, 3), (192, 64), (3, 5, 1, 1), (512, 512, 3, 3)])
def test_pack_weights_is_permute_and_rne_cast(ops, shape, dtype):
    """sg_pack_weights: state_dict weight [Cout, Cin, kh, kw] -> [taps, Cout, Cin] in the operand dtype, bit-identical
    to torch's permute + .to(dtype) (round to nearest even), incl. values that overflow fp16."""
    w = torch.randn(shape, generator=gen(31), dtype=torch.float32)
    w.view(-1)[:4] = torch.tensor([70000.0, -70000.0, 1e-8, 65519.9])
    got = ops.pack_weights(w.to(DEV), dtype)
    cout, cin = shape[:2]
    want = w.reshape(cout, cin, -1).permute(2, 0, 1).contiguous().to(dtype)
    assert got.shape == want.shape and got.dtype == dtype
    assert torch.equal(got.cpu().view(torch.int16 if dtype != torch.float32 else torch.int32),
                       want.view(torch.int16 if dtype != torch.float32 else torch.int32))
