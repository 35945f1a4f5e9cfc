"""The thin C-ABI torch custom-op layer: every kernel entry point of include/sgb200.h is a `torch.ops.sgb200.*` operator.

Each operator is registered with `torch.library` for the CUDA dispatch key ONLY: its implementation checks shapes and
dtypes and launches the C-ABI function (ctypes, `_cabi`) on torch's current stream; outputs are caller-allocated tensors
declared as mutated arguments in the schema (no allocation, no synchronisation, capturable in a CUDA graph).  There is no
CPU or composite kernel, so calling an operator with CPU tensors fails in the dispatcher ("could not run sgb200::... with
arguments from the 'CPU' backend"): nothing here computes on the host and nothing falls back to PyTorch ops.

The module-level functions are the same operators under their Python names; a few of them first flatten convenience
arguments ((hi, lo) operand pairs of the split-TF32 engine, the `outc` triple, uint64 seeds) or allocate a result.
"""
from __future__ import annotations

import functools

import torch

from . import _cabi
from ._cabi import SG_ENGINE_SIMT, SG_ENGINE_TC, IgemmArgs, check, dtype_code, ptr, stream_ptr

HEADS = 4  # nn.MultiheadAttention(channels, 4) -- /root/reference/src/diff_modules.py:56
FUSED_TOKEN_C = (64, 128)  # channel counts of the fused SelfAttention head / tail kernels

_TORCH_LIB = torch.library.Library("sgb200", "DEF")


def _torch_op(schema: str):
    """Register the decorated function (named _<op>) as the CUDA implementation of torch.ops.sgb200.<op> and return the
    dispatcher entry point."""

    def deco(fn):
        op_name = fn.__name__.lstrip("_")
        _TORCH_LIB.define(op_name + schema)
        _TORCH_LIB.impl(op_name, fn, "CUDA")
        packet = getattr(torch.ops.sgb200, op_name)

        @functools.wraps(fn)
        def call(*a, **kw):
            return packet(*a, **kw)

        call.__name__ = op_name
        call.op = packet
        return call

    return deco


def _lib():
    return _cabi.load()


def _f32(t, name):
    if t is not None and (t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA fp32 tensor")
    return t


def _host_f32(t, name):
    """Small weights that travel by value as kernel launch parameters: a contiguous fp32 HOST tensor."""
    if t.dtype != torch.float32 or t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous fp32 HOST tensor (it is passed by value as a launch parameter)")
    return t


def _engine_of(dt: torch.dtype) -> int:
    return SG_ENGINE_SIMT if dt == torch.float32 else SG_ENGINE_TC


def _seed_i64(seed: int) -> int:
    """uint64 Philox seed -> the int64 a dispatcher `int` carries (two's complement); _seed_u64 undoes it."""
    s = int(seed) & (2**64 - 1)
    return s - 2**64 if s >= 2**63 else s


def _seed_u64(seed: int) -> int:
    return int(seed) & (2**64 - 1)


# ---------------------------------------------------------------------------------------------------------------------
# K5 / inc / K1
# ---------------------------------------------------------------------------------------------------------------------
@_torch_op("(Tensor? t, Tensor? step, Tensor? y, Tensor inv_freq, Tensor? label, Tensor w_emb, Tensor b_emb, "
           "Tensor(a!) temb, Tensor(b!) emb) -> ()")
def _time_embed(t, step, y, inv_freq, label, w_emb, b_emb, temb, emb):
    """K5.  t fp32 [rows] or None (then step int32[1] is used); y int64 [rows] or None."""
    rows = temb.shape[0]
    num_classes = 0 if label is None else label.shape[0]
    check(_lib().sg_time_embed(ptr(t), ptr(step), ptr(y), ptr(inv_freq), ptr(label), num_classes, ptr(w_emb),
                               ptr(b_emb), w_emb.shape[0], rows, ptr(temb), ptr(emb), stream_ptr()), "sg_time_embed")


time_embed = _time_embed


def conv_in_partials(S: int) -> int:
    return _lib().sg_conv_in_partials(S)


@_torch_op("(Tensor x, Tensor w, Tensor(a!) raw, Tensor(b!) partials) -> ()")
def _conv_in(x, w, raw, partials):
    """inc.double_conv.0.  x fp32 NCHW [n_src,c,S,S]; w fp32 HOST [64,c,3,3]; raw fp32 or fp16 [rows,S,S,64];
    partials fp32 [rows,P,2]."""
    n_src, c_in, S, _ = x.shape
    rows = raw.shape[0]
    if raw.dtype not in (torch.float32, torch.float16):
        raise ValueError("conv_in: raw must be fp32 or fp16")
    if tuple(w.shape) != (64, c_in, 3, 3):
        raise ValueError(f"conv_in: w must be [64, {c_in}, 3, 3], got {tuple(w.shape)}")
    check(_lib().sg_conv_in(ptr(_f32(x, "x")), n_src, c_in, S, ptr(_host_f32(w, "w")), rows, ptr(raw),
                            dtype_code(raw.dtype), ptr(partials), stream_ptr()), "sg_conv_in")


conv_in = _conv_in


def igemm_partials(engine: int, H: int, W: int, Cout: int) -> int:
    return _lib().sg_igemm_partials(engine, H, W, Cout)


@_torch_op("(Tensor a, Tensor? a_lo, Tensor w, Tensor? w_lo, int rows, int H, int W, Tensor? bias, Tensor? residual, "
           "Tensor(a!)? out_f32, Tensor(b!)? out_act, Tensor(c!)? partials, int act) -> ()")
def _igemm(a, a_lo, w, w_lo, rows, H, W, bias, residual, out_f32, out_act, partials, act):
    """K1.  a: act [rows,H,W,Cin] (any view with that many elements); w: act [taps,Cout,Cin]; a_lo / w_lo: the lo parts
    of the split-TF32 engine (fp32 operands) or None.  act: sg_act."""
    if (a_lo is None) != (w_lo is None):
        raise ValueError("igemm: the split-tf32 engine needs lo parts for both the activation and the weight")
    if a_lo is not None:
        for t in (a, a_lo, w, w_lo):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise ValueError("igemm: split operands must be contiguous CUDA fp32 tensors")
        if a_lo.shape != a.shape or w_lo.shape != w.shape:
            raise ValueError("igemm: hi / lo shape mismatch")
    taps, Cout, Cin = w.shape
    if a.dtype != w.dtype:
        raise ValueError("activation / weight dtype mismatch")
    if a.numel() != rows * H * W * Cin:
        raise ValueError(f"igemm: a has {a.numel()} elements, expected {rows}x{H}x{W}x{Cin}")
    M = rows * H * W
    for t, nm in ((residual, "residual"), (out_f32, "out_f32")):
        if t is not None and (t.dtype != torch.float32 or t.numel() != M * Cout):
            raise ValueError(f"igemm: {nm} must be fp32 with {M}x{Cout} elements")
    out_dtype = 0
    if out_act is not None:
        if out_act.numel() != M * Cout:
            raise ValueError("igemm: out_act must have M x Cout elements")
        if out_act.dtype != a.dtype:
            if a.dtype == torch.float32 or out_act.dtype not in (torch.float16, torch.bfloat16):
                raise ValueError("igemm: out_act must have the activation dtype (or be a 16-bit tensor in the tensor-core modes)")
            out_dtype = dtype_code(out_act.dtype)
    engine = SG_ENGINE_TC if a_lo is not None else _engine_of(a.dtype)
    args = IgemmArgs(ptr(a), ptr(w), ptr(bias), ptr(residual), ptr(out_f32), ptr(out_act), ptr(partials), ptr(a_lo),
                     ptr(w_lo), rows, H, W, Cin, Cout, taps, int(act), engine, dtype_code(a.dtype), out_dtype)
    check(_lib().sg_igemm(args, stream_ptr()), "sg_igemm")


def igemm(a, w, *, rows, H, W, bias=None, residual=None, out_f32=None, out_act=None, partials=None, gelu=False,
          relu_post=False):
    """K1 implicit GEMM (3x3 conv for taps=9, Linear for taps=1) = torch.ops.sgb200.igemm.  The fp32-accurate tensor-core
    engine takes `a` and `w` as (hi, lo) pairs of fp32 tensors (split_tf32); a single fp32 tensor selects the CUDA-core
    engine, 16-bit tensors the tcgen05 kind::f16 engine."""
    a_lo = w_lo = None
    if isinstance(a, tuple) or isinstance(w, tuple):
        if not (isinstance(a, tuple) and isinstance(w, tuple)):
            raise ValueError("igemm: the split-tf32 engine needs (hi, lo) pairs for both the activation and the weight")
        (a, a_lo), (w, w_lo) = a, w
    _igemm(a, a_lo, w, w_lo, rows, H, W, bias, residual, out_f32, out_act, partials,
           _cabi.SG_ACT_GELU if gelu else (_cabi.SG_ACT_RELU_POST if relu_post else _cabi.SG_ACT_NONE))


# ---------------------------------------------------------------------------------------------------------------------
# K2 / K3
# ---------------------------------------------------------------------------------------------------------------------
@_torch_op("(Tensor raw, Tensor partials, Tensor gamma, Tensor beta, *, int mode, Tensor? residual=None, Tensor? emb=None, "
           "Tensor(a!)? out_f32=None, Tensor(b!)? out_act=None, Tensor(c!)? range_flag=None) -> ()")
def _gn_apply(raw, partials, gamma, beta, *, mode, residual=None, emb=None, out_f32=None, out_act=None, range_flag=None):
    """K2.  raw fp32 or fp16 [rows,HW,C] (any shape with rows first, C last); emb: fp32 view [rows, C] (row stride kept).
    range_flag: int32[1] set to 1 by the kernel when an fp16 raw row left fp16's safe range (see sgb200.h)."""
    raw_rows, C = raw.shape[0], raw.shape[-1]
    HW = raw.numel() // (raw_rows * C)
    P = partials.shape[1]
    out_any = out_f32 if out_f32 is not None else out_act
    rows = out_any.shape[0] if out_any is not None else raw_rows  # > raw_rows: row r normalises raw row r % raw_rows
    if rows % raw_rows or partials.shape[0] != raw_rows:
        raise ValueError("gn_apply: output rows must be a multiple of the raw rows; partials must have the raw rows")
    emb_stride = 0
    if emb is not None:
        if emb.dtype != torch.float32 or emb.shape != (rows, C) or emb.stride(1) != 1:
            raise ValueError("gn_apply: emb must be an fp32 [rows, C] view with unit inner stride")
        emb_stride = emb.stride(0)
    adt = dtype_code(out_act.dtype) if out_act is not None else 0
    if raw.dtype not in (torch.float32, torch.float16):
        raise ValueError("gn_apply: raw must be fp32 or fp16")
    check(_lib().sg_gn_apply(ptr(raw), dtype_code(raw.dtype), ptr(partials), P, ptr(gamma), ptr(beta), rows, raw_rows, HW, C, mode,
                             ptr(_f32(residual, "residual")), ptr(emb), emb_stride, ptr(_f32(out_f32, "out_f32")),
                             ptr(out_act), adt, ptr(range_flag), stream_ptr()), "sg_gn_apply")


gn_apply = _gn_apply


@_torch_op("(Tensor raw, Tensor partials, Tensor gamma, Tensor beta, Tensor x, Tensor skip, Tensor(a!) out_act, "
           "Tensor(b!)? range_flag=None) -> ()")
def _gn_apply_vcat(raw, partials, gamma, beta, x, skip, out_act, range_flag=None):
    """GELU(GroupNorm(raw) + cat([skip, upsample2x(x)])) -> 16 bit (sg_gn_apply_vcat).  raw fp16 [rows,2h,2w,Cs+Cx];
    x fp32 [rows,h,w,Cx]; skip fp32 [skip_rows,2h,2w,Cs]."""
    rows, h, w, Cx = x.shape
    Cs = skip.shape[-1]
    if raw.dtype != torch.float16 or tuple(raw.shape) != (rows, 2 * h, 2 * w, Cs + Cx):
        raise ValueError("gn_apply_vcat: raw must be fp16 [rows, 2h, 2w, Cs+Cx]")
    if tuple(skip.shape[1:3]) != (2 * h, 2 * w) or rows % skip.shape[0]:
        raise ValueError("gn_apply_vcat: skip must be [skip_rows, 2h, 2w, Cs] with rows a multiple of skip_rows")
    check(_lib().sg_gn_apply_vcat(ptr(raw), ptr(partials), partials.shape[1], ptr(gamma), ptr(beta), rows,
                                  ptr(_f32(x, "x")), ptr(_f32(skip, "skip")), skip.shape[0], h, w, Cx, Cs, ptr(out_act),
                                  dtype_code(out_act.dtype), ptr(range_flag), stream_ptr()), "sg_gn_apply_vcat")


gn_apply_vcat = _gn_apply_vcat


@_torch_op("(Tensor x, *, Tensor(a!)? out_f32=None, Tensor(b!)? out_act=None) -> ()")
def _maxpool2(x, *, out_f32=None, out_act=None):
    """K3a.  x fp32 [rows,H,W,C]."""
    rows, H, W, Cc = x.shape
    adt = dtype_code(out_act.dtype) if out_act is not None else 0
    check(_lib().sg_maxpool2(ptr(_f32(x, "x")), rows, H, W, Cc, ptr(_f32(out_f32, "out_f32")), ptr(out_act), adt,
                             stream_ptr()), "sg_maxpool2")


maxpool2 = _maxpool2


@_torch_op("(Tensor x, Tensor skip, *, Tensor(a!)? out_f32=None, Tensor(b!)? out_act=None) -> ()")
def _upsample_cat(x, skip, *, out_f32=None, out_act=None):
    """K3b.  x fp32 [rows,h,w,Cx], skip fp32 [skip_rows,2h,2w,Cs] -> [rows,2h,2w,Cs+Cx]; row r reads skip row
    r % skip_rows (skip_rows == rows, or rows/2 for the shared label-independent prefix)."""
    rows, h, w, Cx = x.shape
    Cs = skip.shape[-1]
    skip_rows = skip.shape[0]
    if tuple(skip.shape[1:3]) != (2 * h, 2 * w) or rows % skip_rows:
        raise ValueError("upsample_cat: skip must be [skip_rows, 2h, 2w, Cs] with rows a multiple of skip_rows")
    adt = dtype_code(out_act.dtype) if out_act is not None else 0
    check(_lib().sg_upsample_cat(ptr(_f32(x, "x")), ptr(_f32(skip, "skip")), rows, skip_rows, h, w, Cx, Cs,
                                 ptr(_f32(out_f32, "out_f32")), ptr(out_act), adt, stream_ptr()), "sg_upsample_cat")


upsample_cat = _upsample_cat


@_torch_op("(Tensor x, Tensor gamma, Tensor beta, Tensor(a!) out) -> ()")
def _layernorm(x, gamma, beta, out):
    """LayerNorm over the last dim.  x fp32 [..., C] -> out (fp32 / bf16 / fp16) of the same shape."""
    Cc = x.shape[-1]
    M = x.numel() // Cc
    check(_lib().sg_layernorm(ptr(_f32(x, "x")), ptr(gamma), ptr(beta), M, Cc, ptr(out), dtype_code(out.dtype),
                              stream_ptr()), "sg_layernorm")


layernorm = _layernorm


# ---------------------------------------------------------------------------------------------------------------------
# DiffusionVAE decode tail
# ---------------------------------------------------------------------------------------------------------------------
@_torch_op("(Tensor x, Tensor codebook, Tensor(a!) quantized, Tensor(b!)? indices=None, *, bool clamp=True) -> ()")
def _vq_quantize(x, codebook, quantized, indices=None, *, clamp=True):
    """[clamp(-1,1) +] nearest-codeword quantisation of groups of 4 consecutive fp32 values (VQEmbeddingEMA.forward)."""
    check(_lib().sg_vq_quantize(ptr(_f32(x, "x")), x.numel(), ptr(_f32(codebook, "codebook")), codebook.shape[0],
                                1 if clamp else 0, ptr(_f32(quantized, "quantized")), ptr(indices), stream_ptr()),
          "sg_vq_quantize")


vq_quantize = _vq_quantize


@_torch_op("(Tensor z, Tensor w, Tensor b, *, Tensor(a!)? out_f32=None, Tensor(b!)? out_act=None) -> ()")
def _dec_in_proj(z, w, b, *, out_f32=None, out_act=None):
    """Decoder.in_proj: z fp32 NCHW [n,4,S,S] -> NHWC [n,S,S,Cout]."""
    n, _, S, _ = z.shape
    adt = dtype_code(out_act.dtype) if out_act is not None else 0
    check(_lib().sg_dec_in_proj(ptr(_f32(z, "z")), ptr(_f32(w, "w")), ptr(b), n, S, w.shape[0], ptr(out_f32), ptr(out_act),
                                adt, stream_ptr()), "sg_dec_in_proj")


dec_in_proj = _dec_in_proj


@_torch_op("(Tensor t, Tensor w2, Tensor b2, *, int n, int S, Tensor(a!)? out_u8=None, Tensor(b!)? out_f32=None) -> ()")
def _tconv2_u8(t, w2, b2, *, n, S, out_u8=None, out_f32=None):
    """Decoder.strided_t_conv_2 + uint8 image tail on the un-shuffled first transposed conv t [2, n*S*S, 2*C]."""
    Cc = w2.shape[0]
    check(_lib().sg_tconv2_u8(ptr(t), dtype_code(t.dtype), n, S, Cc, ptr(_f32(w2, "w2")), ptr(b2), ptr(out_u8),
                              ptr(out_f32), stream_ptr()), "sg_tconv2_u8")


tconv2_u8 = _tconv2_u8


# ---------------------------------------------------------------------------------------------------------------------
# SelfAttention
# ---------------------------------------------------------------------------------------------------------------------
@_torch_op("(Tensor x, Tensor ln_g, Tensor ln_b, Tensor w_in, Tensor b_in, Tensor(a!) qkv) -> ()")
def _ln_inproj(x, ln_g, ln_b, w_in, b_in, qkv):
    """Fused LayerNorm + in_proj (tcgen05).  x fp32 [M, C]; w_in 16-bit [.., 3C, C]; qkv 16-bit [M, 3C]."""
    Cc = x.shape[-1]
    M = x.numel() // Cc
    check(_lib().sg_ln_inproj(ptr(_f32(x, "x")), ptr(ln_g), ptr(ln_b), ptr(w_in), ptr(b_in), M, Cc, ptr(qkv),
                              dtype_code(qkv.dtype), stream_ptr()), "sg_ln_inproj")


ln_inproj = _ln_inproj


@_torch_op("(Tensor att, Tensor x, Tensor wo, Tensor bo, Tensor ln_g, Tensor ln_b, Tensor w1, Tensor b1, Tensor w2, "
           "Tensor b2, Tensor(a!)? out, Tensor? outc_w, Tensor? outc_b, Tensor(b!)? eps) -> ()")
def _attn_tail(att, x, wo, bo, ln_g, ln_b, w1, b1, w2, b2, out, outc_w, outc_b, eps):
    Cc = x.shape[-1]
    M = x.numel() // Cc
    if eps is not None:
        c_out, HW = eps.shape[1], eps.shape[2] * eps.shape[3]
        if eps.numel() != (M // HW) * c_out * HW or M % HW:
            raise ValueError("attn_tail: eps does not match the token count")
        check(_lib().sg_attn_tail_outc(ptr(att), ptr(_f32(x, "x")), ptr(wo), ptr(bo), ptr(ln_g), ptr(ln_b), ptr(w1),
                                       ptr(b1), ptr(w2), ptr(b2), M, Cc, ptr(_f32(out, "out")), ptr(_host_f32(outc_w, "outc_w")),
                                       ptr(_host_f32(outc_b, "outc_b")), c_out, HW, ptr(_f32(eps, "eps")), dtype_code(att.dtype),
                                       stream_ptr()), "sg_attn_tail_outc")
        return
    check(_lib().sg_attn_tail(ptr(att), ptr(_f32(x, "x")), ptr(wo), ptr(bo), ptr(ln_g), ptr(ln_b), ptr(w1), ptr(b1),
                              ptr(w2), ptr(b2), M, Cc, ptr(_f32(out, "out")), dtype_code(att.dtype), stream_ptr()),
          "sg_attn_tail")


def attn_tail(att, x, wo, bo, ln_g, ln_b, w1, b1, w2, b2, out, *, outc=None):
    """Fused out_proj + residual + LayerNorm + FFN + residual (tcgen05) = torch.ops.sgb200.attn_tail.  att 16-bit [M, C];
    x, out fp32 [M, C].  outc = (w fp32 HOST [c_out, C], b fp32 HOST [c_out], eps fp32 NCHW [rows, c_out, S, S]) also
    applies the model's 1x1 output conv to the block output (sg_attn_tail_outc); `out` may then be None."""
    w, b, eps = outc if outc is not None else (None, None, None)
    _attn_tail(att, x, wo, bo, ln_g, ln_b, w1, b1, w2, b2, out, w, b, eps)


@_torch_op("(Tensor qkv, Tensor(a!) out, *, int rows, int L, int C, int engine) -> ()")
def _attention(qkv, out, *, rows, L, C, engine):
    if qkv.numel() != rows * L * 3 * C or out.numel() != rows * L * C:
        raise ValueError("attention: bad qkv / out size")
    check(_lib().sg_attention(ptr(qkv), ptr(out), rows, L, C, HEADS, engine, dtype_code(out.dtype), stream_ptr()),
          "sg_attention")


def attention(qkv, out, *, rows, L, C, engine=None):
    """K4.  qkv [rows*L, 3C]; out [rows*L, C].  SIMT engine: qkv fp32, out fp32/16-bit.  TC: both 16-bit."""
    _attention(qkv, out, rows=rows, L=L, C=C, engine=_engine_of(qkv.dtype) if engine is None else engine)


@_torch_op("(Tensor qkv, Tensor(a!) qk_hi, Tensor(b!) qk_lo, Tensor(c!) vt_hi, Tensor(d!) vt_lo, *, int rows, int L, int C) -> ()")
def _attn_prep_tf32(qkv, qk_hi, qk_lo, vt_hi, vt_lo, *, rows, L, C):
    """Operand form of the split-TF32 attention core: qkv fp32 [rows*L, 3C] -> (hi, lo) of q | k [rows*L, 2C] and of V
    transposed per (batch row, head) [rows*C, L]."""
    M = rows * L
    if qkv.numel() != M * 3 * C or qk_hi.numel() != M * 2 * C or qk_lo.numel() != M * 2 * C or vt_hi.numel() != M * C \
            or vt_lo.numel() != M * C:
        raise ValueError("attn_prep_tf32: bad buffer sizes")
    check(_lib().sg_attn_prep_tf32(ptr(_f32(qkv, "qkv")), ptr(_f32(qk_hi, "qk_hi")), ptr(_f32(qk_lo, "qk_lo")),
                                   ptr(_f32(vt_hi, "vt_hi")), ptr(_f32(vt_lo, "vt_lo")), rows, L, C, HEADS, stream_ptr()),
          "sg_attn_prep_tf32")


attn_prep_tf32 = _attn_prep_tf32


@_torch_op("(Tensor qk_hi, Tensor qk_lo, Tensor vt_hi, Tensor vt_lo, Tensor(a!) out, *, int rows, int L, int C) -> ()")
def _attention_tf32(qk_hi, qk_lo, vt_hi, vt_lo, out, *, rows, L, C):
    """K4 of the fp32-accurate tensor-core engine (L >= 128); operands from attn_prep_tf32, out fp32 [rows*L, C]."""
    M = rows * L
    if qk_hi.numel() != M * 2 * C or qk_lo.numel() != M * 2 * C or vt_hi.numel() != M * C or vt_lo.numel() != M * C \
            or out.numel() != M * C:
        raise ValueError("attention_tf32: bad buffer sizes")
    check(_lib().sg_attention_tf32(ptr(_f32(qk_hi, "qk_hi")), ptr(_f32(qk_lo, "qk_lo")), ptr(_f32(vt_hi, "vt_hi")),
                                   ptr(_f32(vt_lo, "vt_lo")), ptr(_f32(out, "out")), rows, L, C, HEADS, stream_ptr()),
          "sg_attention_tf32")


attention_tf32 = _attention_tf32


@_torch_op("(Tensor x, Tensor w, Tensor b, Tensor(a!) eps) -> ()")
def _conv_out(x, w, b, eps):
    """outc.  x fp32 [rows,HW,64] -> eps fp32 NCHW [rows,c_out,S,S]."""
    rows, c_out = eps.shape[0], eps.shape[1]
    HW = eps.shape[2] * eps.shape[3]
    check(_lib().sg_conv_out(ptr(_f32(x, "x")), ptr(w), ptr(b), rows, HW, c_out, ptr(eps), stream_ptr()), "sg_conv_out")


conv_out = _conv_out


# ---------------------------------------------------------------------------------------------------------------------
# sampler
# ---------------------------------------------------------------------------------------------------------------------
@_torch_op("(Tensor(a!) x, Tensor eps, Tensor coef, Tensor step, *, float cfg_scale, Tensor? noise=None, int seed=0, "
           "int sample_base=0) -> ()")
def _cfg_update(x, eps, coef, step, *, cfg_scale, noise=None, seed=0, sample_base=0):
    n = x.shape[0]
    E = x.numel() // n
    T = coef.shape[0]
    need = 2 * n if cfg_scale > 0 else n
    if eps.shape[0] < need:
        raise ValueError(f"cfg_update: eps has {eps.shape[0]} rows, needs {need}")
    if noise is not None and (noise.dtype != torch.float32 or noise.numel() != (T - 1) * n * E):
        raise ValueError("cfg_update: injected noise must be fp32 [T-1, n, c, S, S]")
    check(_lib().sg_cfg_update(ptr(_f32(x, "x")), ptr(_f32(eps, "eps")), n, E, float(cfg_scale), ptr(coef), T,
                               ptr(step), ptr(noise), _seed_u64(seed), int(sample_base), stream_ptr()),
          "sg_cfg_update")


def cfg_update(x, eps, coef, step, *, cfg_scale, noise=None, seed=0, sample_base=0):
    """K6.  x fp32 [n,c,S,S] in/out; eps fp32 [2n or n, c,S,S]; coef fp32 [T,3]; step int32[1]; seed: any uint64."""
    _cfg_update(x, eps, coef, step, cfg_scale=float(cfg_scale), noise=noise, seed=_seed_i64(seed),
                sample_base=int(sample_base))


@_torch_op("(Tensor(a!) step) -> ()")
def _step_advance(step):
    check(_lib().sg_step_advance(ptr(step), stream_ptr()), "sg_step_advance")


step_advance = _step_advance


@_torch_op("(Tensor(a!) x, *, int seed, int sample_base, int step_tag) -> ()")
def _philox_normal(x, *, seed, sample_base, step_tag):
    n = x.shape[0]
    check(_lib().sg_philox_normal(ptr(_f32(x, "x")), n, x.numel() // n, _seed_u64(seed), int(sample_base),
                                  int(step_tag), stream_ptr()), "sg_philox_normal")


def philox_normal(x, *, seed, sample_base, step_tag):
    _philox_normal(x, seed=_seed_i64(seed), sample_base=int(sample_base), step_tag=int(step_tag))


@_torch_op("(Tensor x, Tensor(a!) out, *, bool wrap=False) -> ()")
def _to_uint8(x, out, *, wrap=False):
    if out.dtype != torch.uint8 or not out.is_contiguous() or out.numel() != x.numel():
        raise ValueError("to_uint8: out must be a contiguous uint8 tensor of x's size")
    fn = _lib().sg_to_uint8_wrap if wrap else _lib().sg_to_uint8
    check(fn(ptr(_f32(x, "x")), x.numel(), ptr(out), stream_ptr()), "sg_to_uint8_wrap" if wrap else "sg_to_uint8")


def to_uint8(x, out):
    """K8: (clamp(x,-1,1)+1)/2*255 -> truncating uint8 cast (:440-441)."""
    _to_uint8(x, out, wrap=False)


def to_uint8_wrap(x, out=None):
    """uint8((x + 1) / 2 * 255) without a clamp (the cast of the reference's trajectory dumps, :672-675)."""
    if out is None:
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _to_uint8(x, out, wrap=True)
    return out


def set_pdl(mode: int):
    """Programmatic dependent launch mode of this thread's following launches (0 off, 1 all, 2 small grids)."""
    check(_lib().sg_set_pdl(int(mode)), "sg_set_pdl")


# ---------------------------------------------------------------------------------------------------------------------
# operand preparation
# ---------------------------------------------------------------------------------------------------------------------
@_torch_op("(Tensor x, Tensor(a!) hi, Tensor(b!) lo) -> ()")
def _split_tf32(x, hi, lo):
    if hi.numel() != x.numel() or lo.numel() != x.numel():
        raise ValueError("split_tf32: hi / lo must have x's size")
    check(_lib().sg_split_tf32(ptr(_f32(x, "x")), ptr(_f32(hi, "hi")), ptr(_f32(lo, "lo")), x.numel(), stream_ptr()),
          "sg_split_tf32")


@_torch_op("(Tensor x, Tensor(a!) lo) -> ()")
def _split_tf32_lo(x, lo):
    if lo.numel() != x.numel():
        raise ValueError("split_tf32_lo: lo must have x's size")
    check(_lib().sg_split_tf32(ptr(_f32(x, "x")), None, ptr(_f32(lo, "lo")), x.numel(), stream_ptr()), "sg_split_tf32")


def split_tf32_lo(x, lo=None):
    """Activation operand form of the fp32-accurate tensor-core engine: x itself is the high part (the tensor core reads
    its top 19 bits), returns lo = tf32(x - trunc19(x)).  gn_apply / maxpool2 / upsample_cat write the same lo tensor
    themselves when given an fp32 `out_act`."""
    lo = torch.empty_like(x) if lo is None else lo
    _split_tf32_lo(x, lo)
    return lo


def split_tf32(x, hi=None, lo=None):
    """x fp32 -> (hi, lo): hi = tf32(x), lo = tf32(x - hi) -- the operand form of the fp32-accurate tensor-core engine."""
    hi = torch.empty_like(x) if hi is None else hi
    lo = torch.empty_like(x) if lo is None else lo
    _split_tf32(x, hi, lo)
    return hi, lo


@_torch_op("(Tensor w, Tensor(a!) out) -> ()")
def _pack_weights(w, out):
    cout, cin = w.shape[0], w.shape[1]
    taps = w.numel() // (cout * cin)
    if tuple(out.shape) != (taps, cout, cin):
        raise ValueError("pack_weights: out must be [taps, Cout, Cin]")
    check(_lib().sg_pack_weights(ptr(_f32(w, "w")), cout, cin, taps, ptr(out), dtype_code(out.dtype), stream_ptr()),
          "sg_pack_weights")


def pack_weights(w, dtype):
    """fp32 device weight [Cout, Cin, *kernel] -> [taps, Cout, Cin] of `dtype` (the B operand of sg_igemm)."""
    cout, cin = w.shape[0], w.shape[1]
    out = torch.empty((w.numel() // (cout * cin), cout, cin), dtype=dtype, device=w.device)
    _pack_weights(w, out)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# forward-only training helpers
# ---------------------------------------------------------------------------------------------------------------------
@_torch_op("(Tensor x, Tensor t, Tensor alpha_hat, Tensor? eps_in, int seed, int sample_base, Tensor(a!) x_t, "
           "Tensor(b!)? eps_out) -> ()")
def _noise_images(x, t, alpha_hat, eps_in, seed, sample_base, x_t, eps_out):
    n = x.shape[0]
    E = x.numel() // max(n, 1)
    check(_lib().sg_noise_images(ptr(_f32(x, "x")), ptr(t), ptr(_f32(alpha_hat, "alpha_hat")), alpha_hat.numel(), n, E,
                                 ptr(_f32(eps_in, "eps")), _seed_u64(seed), int(sample_base), ptr(x_t), ptr(eps_out),
                                 stream_ptr()), "sg_noise_images")


def noise_images(x, t, alpha_hat, *, eps=None, seed=0, sample_base=0):
    """Diffusion.noise_images (:404-409): returns (x_t, eps).  x fp32 [n, ...]; t int64 [n]; alpha_hat fp32 [T]."""
    x = _f32(x, "x")
    n = x.shape[0]
    if t.dtype != torch.int64 or not t.is_cuda or t.numel() != n:
        raise ValueError("noise_images: t must be a CUDA int64 tensor with one timestep per sample")
    T = alpha_hat.numel()
    if n > 0 and (int(t.min()) < -T or int(t.max()) >= T):
        raise IndexError(f"index out of range: timesteps must lie in [-{T}, {T})")  # as alpha_hat[t] raises in the reference
    if eps is not None and (eps.shape != x.shape):
        raise ValueError("noise_images: eps must have x's shape")
    x_t = torch.empty_like(x)
    eps_out = torch.empty_like(x) if eps is None else None
    _noise_images(x, t.contiguous(), alpha_hat, eps, _seed_i64(seed), int(sample_base), x_t, eps_out)
    return x_t, (eps if eps is not None else eps_out)


@_torch_op("(Tensor(a!) ma, Tensor cur, float beta) -> ()")
def _ema_update(ma, cur, beta):
    if ma.shape != cur.shape:
        raise ValueError("ema_update: shape mismatch")
    check(_lib().sg_ema_update(ptr(_f32(ma, "ma")), ptr(_f32(cur, "cur")), ma.numel(), float(beta), float(1 - beta),
                               stream_ptr()), "sg_ema_update")


def ema_update(ma, cur, beta):
    """EMA.update_average (:37-40) in place on ma: ma * beta + (1 - beta) * cur; returns ma."""
    _ema_update(ma, cur, float(beta))
    return ma


@_torch_op("(Tensor a, Tensor b, Tensor(a!) scratch, Tensor(b!) out) -> ()")
def _mse(a, b, scratch, out):
    if a.shape != b.shape or a.numel() == 0:
        raise ValueError("mse: shape mismatch / empty input")
    check(_lib().sg_mse(ptr(_f32(a, "a")), ptr(_f32(b, "b")), a.numel(), ptr(scratch), ptr(out), stream_ptr()), "sg_mse")


def mse(a, b):
    """nn.MSELoss() (:478): 0-dim fp32 tensor."""
    scratch = torch.empty(_lib().sg_mse_scratch_doubles(), dtype=torch.float64, device=a.device)
    out = torch.empty((), dtype=torch.float32, device=a.device)
    _mse(a, b, scratch, out)
    return out
