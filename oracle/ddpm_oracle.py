"""CPU restatement of the reference CFG-DDPM sampling path -- TEST INFRASTRUCTURE ONLY.

A state_dict-driven, functional restatement (no nn.Module, explicit attention) of
  * UNet_conditional.forward     /root/reference/src/diff_modules.py:204-217, :175-196
  * DoubleConv / Down / Up       :75-93 / :96-113 / :116-136
  * SelfAttention                :52-72   (torch nn.MultiheadAttention, 4 heads, batch_first)
  * pos_encoding                 :168-173
  * Diffusion schedule + sample  :387-389, :398-399, :411-442
The arithmetic of the reference lives in un-vendored PyTorch (torch pinned 2.12.1 in
/root/reference/uv.lock; this image has 2.11.0): conv/linear/norm primitives are therefore
called from torch.nn.functional on CPU tensors, in fp32 (the reference's dtype, and the
honest CPU baseline: same ATen/oneDNN kernels, same thread count) or fp64 (ground truth
for the tolerance budget).  Attention is written out (softmax(QK^T/sqrt(d))V) instead of
calling nn.MultiheadAttention, so the O(L^2) head-averaged weights the reference computes
and throws away (:69, need_weights default) are not computed.

Pinned against the reference itself by tests/golden/ (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

NUM_HEADS = 4  # diff_modules.py:56
EPS = 1e-5  # torch default eps of GroupNorm / LayerNorm
ATTN_QUERY_CHUNK = 4096  # query rows per materialised block of the attention matrix


# --------------------------------------------------------------------------------------
# schedule (diff_modules.py:387-389, :398-399)
# --------------------------------------------------------------------------------------
def noise_schedule(noise_steps=1000, beta_start=1e-4, beta_end=0.02):
    beta = torch.linspace(beta_start, beta_end, noise_steps)
    alpha = 1.0 - beta
    alpha_hat = torch.cumprod(alpha, dim=0)
    return beta, alpha, alpha_hat


def posterior_coefficients(beta, alpha, alpha_hat):
    """c1, c2, c3 of x <- c1*(x - c2*eps) + c3*z, evaluated as the reference writes them (:436-439)."""
    c1 = 1 / torch.sqrt(alpha)
    c2 = (1 - alpha) / (torch.sqrt(1 - alpha_hat))
    c3 = torch.sqrt(beta)
    return c1, c2, c3


# --------------------------------------------------------------------------------------
# model blocks
# --------------------------------------------------------------------------------------
def pos_encoding(t, channels=256, dtype=torch.float32):
    """diff_modules.py:168-173.  t: [n] (any numeric dtype) -> [n, channels].

    The reference always evaluates this in fp32 (`.float()` at :169); for dtype=float64 the
    fp32 result is up-cast so both precisions see the same embedding inputs.
    """
    t = t.reshape(-1, 1)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
    arg = t.repeat(1, channels // 2) * inv_freq
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1).to(dtype)


def _double_conv(sd, p, x, residual=False):
    """diff_modules.py:75-93."""
    h = F.conv2d(x, sd[f"{p}.double_conv.0.weight"], padding=1)
    h = F.group_norm(h, 1, sd[f"{p}.double_conv.1.weight"], sd[f"{p}.double_conv.1.bias"], EPS)
    h = F.gelu(h)
    h = F.conv2d(h, sd[f"{p}.double_conv.3.weight"], padding=1)
    h = F.group_norm(h, 1, sd[f"{p}.double_conv.4.weight"], sd[f"{p}.double_conv.4.bias"], EPS)
    return F.gelu(x + h) if residual else h


def _emb(sd, p, temb):
    """emb_layer = SiLU -> Linear(256, Cout) (:105-108, :126-129), broadcast over H, W."""
    return F.linear(F.silu(temb), sd[f"{p}.emb_layer.1.weight"], sd[f"{p}.emb_layer.1.bias"])[:, :, None, None]


def _down(sd, p, x, temb):
    """diff_modules.py:96-113."""
    x = F.max_pool2d(x, 2)
    x = _double_conv(sd, f"{p}.maxpool_conv.1", x, residual=True)
    x = _double_conv(sd, f"{p}.maxpool_conv.2", x)
    return x + _emb(sd, p, temb)


def _up(sd, p, x, skip, temb):
    """diff_modules.py:116-136 (skip first in the concat, :133)."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = torch.cat([skip, x], dim=1)
    x = _double_conv(sd, f"{p}.conv.0", x, residual=True)
    x = _double_conv(sd, f"{p}.conv.1", x)
    return x + _emb(sd, p, temb)


def _self_attention(sd, p, x):
    """diff_modules.py:52-72 with nn.MultiheadAttention written out."""
    n, c, hh, ww = x.shape
    d = c // NUM_HEADS
    tok = x.reshape(n, c, hh * ww).transpose(1, 2)  # [n, L, C], row-major (h, w) token order
    x_ln = F.layer_norm(tok, (c,), sd[f"{p}.ln.weight"], sd[f"{p}.ln.bias"], EPS)
    qkv = F.linear(x_ln, sd[f"{p}.mha.in_proj_weight"], sd[f"{p}.mha.in_proj_bias"])
    q, k, v = qkv.split(c, dim=-1)

    def heads(z):
        return z.reshape(n, -1, NUM_HEADS, d).transpose(1, 2)  # [n, heads, L, d]

    q, k, v = heads(q) * (d ** -0.5), heads(k), heads(v)
    L = q.shape[2]
    if L <= ATTN_QUERY_CHUNK:
        att = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v  # [n, heads, L, d]
    else:
        # same arithmetic per query row (softmax over ALL keys); only the L x L matrix is built ATTN_QUERY_CHUNK query
        # rows at a time so that L = 16384 / 65536 (256 x 256 spectrograms, :376-378) fits the host memory
        att = torch.cat([torch.softmax(q[:, :, i:i + ATTN_QUERY_CHUNK] @ k.transpose(-1, -2), dim=-1) @ v
                         for i in range(0, L, ATTN_QUERY_CHUNK)], dim=2)
    att = att.transpose(1, 2).reshape(n, -1, c)
    att = F.linear(att, sd[f"{p}.mha.out_proj.weight"], sd[f"{p}.mha.out_proj.bias"])
    a = att + tok  # residual is the pre-LN x (:70)
    f = F.layer_norm(a, (c,), sd[f"{p}.ff_self.0.weight"], sd[f"{p}.ff_self.0.bias"], EPS)
    f = F.linear(f, sd[f"{p}.ff_self.1.weight"], sd[f"{p}.ff_self.1.bias"])
    f = F.gelu(f)
    f = F.linear(f, sd[f"{p}.ff_self.3.weight"], sd[f"{p}.ff_self.3.bias"])
    a = f + a
    return a.transpose(1, 2).reshape(n, c, hh, ww)


def cast_state_dict(sd, dtype):
    return {k: v.to(dtype) for k, v in sd.items()}


@torch.inference_mode()
def unet_forward(sd, x, t, y=None, *, dtype=torch.float32, taps=None):
    """UNet_conditional.forward(x, t, y) (:210-217 -> :175-196).

    sd    : state_dict already in `dtype` (see cast_state_dict)
    x     : [n, c_in, S, S];  t: [n] integer or float;  y: [n] int64 or None
    taps  : optional dict filled with the per-block activations (for per-block parity tests)
    """
    x = x.to(dtype)
    temb = pos_encoding(t, 256, dtype)
    if y is not None:
        temb = temb + sd["label_emb.weight"][y]
    deep = "bot2.double_conv.0.weight" in sd

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    tap("temb", temb)
    x1 = tap("inc", _double_conv(sd, "inc", x))
    x2 = tap("down1", _down(sd, "down1", x1, temb))
    x2 = tap("sa1", _self_attention(sd, "sa1", x2))
    x3 = tap("down2", _down(sd, "down2", x2, temb))
    x3 = tap("sa2", _self_attention(sd, "sa2", x3))
    x4 = tap("down3", _down(sd, "down3", x3, temb))
    x4 = tap("sa3", _self_attention(sd, "sa3", x4))
    x4 = tap("bot1", _double_conv(sd, "bot1", x4))
    if deep:
        x4 = tap("bot2", _double_conv(sd, "bot2", x4))
    x4 = tap("bot3", _double_conv(sd, "bot3", x4))
    u = tap("up1", _up(sd, "up1", x4, x3, temb))
    u = tap("sa4", _self_attention(sd, "sa4", u))
    u = tap("up2", _up(sd, "up2", u, x2, temb))
    u = tap("sa5", _self_attention(sd, "sa5", u))
    u = tap("up3", _up(sd, "up3", u, x1, temb))
    u = tap("sa6", _self_attention(sd, "sa6", u))
    return tap("outc", F.conv2d(u, sd["outc.weight"], sd["outc.bias"]))


# --------------------------------------------------------------------------------------
# sampler (diff_modules.py:411-442)
# --------------------------------------------------------------------------------------
def cfg_combine(eps_c, eps_u, cfg_scale):
    """torch.lerp(eps_u, eps_c, s) (:428)."""
    return torch.lerp(eps_u, eps_c, cfg_scale)


def posterior_update(x, eps, c1, c2, c3, z):
    """x <- 1/sqrt(a) * (x - ((1-a)/sqrt(1-ah)) * eps) + sqrt(b) * z   (:436-439), scalar coefficients."""
    return c1 * (x - c2 * eps) + c3 * z


def to_uint8(x):
    """(:440-441): clamp(-1,1) -> (+1)/2 -> *255 -> truncating cast."""
    x = (x.clamp(-1, 1) + 1) / 2
    return (x * 255).type(torch.uint8)


def draw_reference_noise(seed, n, c, s, noise_steps):
    """The draws Diffusion.sample makes on the CPU generator after set_seed(seed), in order:
    x_T (:418) then one randn_like per i = T-1 .. 2 (:432-433); nothing is drawn at i == 1.
    Returns [T-1, n, c, s, s]; index 0 is x_T and index k >= 1 is the z used at i = T-k."""
    torch.manual_seed(seed)
    out = [torch.randn((n, c, s, s))]
    for _ in range(noise_steps - 2):
        out.append(torch.randn((n, c, s, s)))
    return torch.stack(out)


@torch.inference_mode()
def sample(
    sd,
    labels,
    noise,
    *,
    cfg_scale=3,
    noise_steps=1000,
    beta_start=1e-4,
    beta_end=0.02,
    dtype=torch.float32,
    return_float=False,
    trajectory=None,
    max_iters=None,
):
    """Diffusion.sample(use_ema=False, labels, cfg_scale) with the Gaussian draws injected.

    noise: [T-1, n, c, S, S] as produced by draw_reference_noise.  `trajectory`, if a list,
    receives (i, eps, x_after) for every iteration.  `max_iters` bounds the number of loop
    iterations (used by the bounded CPU-baseline timing; the result is then a partial state).
    """
    n = len(labels)
    beta, alpha, alpha_hat = noise_schedule(noise_steps, beta_start, beta_end)
    c1, c2, c3 = posterior_coefficients(beta, alpha, alpha_hat)
    x = noise[0].to(torch.float32)
    it = 0
    for k, i in enumerate(reversed(range(1, noise_steps))):
        if max_iters is not None and it >= max_iters:
            break
        t = (torch.ones(n) * i).long()
        eps = unet_forward(sd, x, t, labels, dtype=dtype).to(torch.float32)
        if cfg_scale > 0:
            eps_u = unet_forward(sd, x, t, None, dtype=dtype).to(torch.float32)
            eps = cfg_combine(eps, eps_u, cfg_scale)
        z = noise[k + 1].to(torch.float32) if i > 1 else torch.zeros_like(x)
        x = posterior_update(x, eps, c1[i], c2[i], c3[i], z)
        if trajectory is not None:
            trajectory.append((i, eps.clone(), x.clone()))
        it += 1
    if return_float:
        return x
    return to_uint8(x)


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def flops_per_forward(s, c_in=4, c_out=4):
    """Algorithmic FLOPs (2*MAC) of one UNet_conditional forward for one sample at S x S
    (conv 2*Cin*Cout*9*h*w; attention block 12*L*C^2 + 4*L^2*C)."""
    def conv(cin, cout, hw, k=9):
        return 2.0 * cin * cout * k * hw

    def dc(cin, cout, hw, mid=None):
        mid = mid or cout
        return conv(cin, mid, hw) + conv(mid, cout, hw)

    def sa(c, L):
        return 12.0 * L * c * c + 4.0 * L * L * c

    def emb(cout):
        return 2.0 * 256 * cout

    s1, s2, s4, s8 = s * s, (s // 2) ** 2, (s // 4) ** 2, (s // 8) ** 2
    f = dc(c_in, 64, s1)
    f += dc(64, 64, s2) + dc(64, 128, s2) + emb(128) + sa(128, s2)
    f += dc(128, 128, s4) + dc(128, 256, s4) + emb(256) + sa(256, s4)
    f += dc(256, 256, s8) + dc(256, 256, s8) + emb(256) + sa(256, s8)
    f += dc(256, 512, s8) + dc(512, 512, s8) + dc(512, 256, s8)
    f += dc(512, 512, s4) + dc(512, 128, s4, 256) + emb(128) + sa(128, s4)
    f += dc(256, 256, s2) + dc(256, 64, s2, 128) + emb(64) + sa(64, s2)
    f += dc(128, 128, s1) + dc(128, 64, s1, 64) + emb(64) + sa(64, s1)
    f += conv(64, c_out, s1, 1)
    return f
