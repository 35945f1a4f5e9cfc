"""Host-side execution plan of UNet_conditional.forward on the sgb200 kernels.

`PackedWeights` repacks a reference-schema state_dict (183 tensors, /root/reference/src/diff_modules.py:75-217)
once into kernel layouts; `UNetPlan` pre-allocates every activation buffer for a fixed (rows, S) geometry and
records the launch sequence as a flat list of closures, so that running the plan performs no allocation, no
host<->device traffic and no synchronisation -- which is what makes it capturable in a CUDA graph.

Data layout in HBM: activations are channels-last [rows, H, W, C] so that (a) SelfAttention's
NCHW -> [n, L, C] transpose (:66) is free, (b) a 3x3 tap is a rectangular TMA box and (c) the implicit GEMM's
K dimension (channels) is contiguous.  The residual stream (block outputs, skips) stays fp32; in the
tensor-core modes every GEMM operand is additionally written as a 16-bit copy by the producing kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _cabi, ops
from ._cabi import SG_ENGINE_SIMT, SG_ENGINE_TC

# compute modes: "bf16" / "f16" = tcgen05 kind::f16 engine (16-bit operands); "fp32" = the fp32-accurate tensor-core engine
# (split-TF32 operands, three kind::tf32 MMAs per product: meets the 1e-4 bar of the reference's fp32 arithmetic at ~1/6 of
# the 16-bit tensor rate); "fp32_simt" = CUDA-core kernels, kept as the independent comparator of the parity tests
MODES = {"fp32": torch.float32, "fp32_simt": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


@dataclass(frozen=True)
class PlanOptions:
    """Structural choices of a UNetPlan.  The defaults are the product path; the alternatives exist because other modes
    need them (the un-fused SelfAttention launches serve C = 256 and the fp32 engines; fp32 raw tensors are the fallback
    of the fp16-range guard) and the parity tests check that every combination gives the same answer."""
    raw16: bool = True           # 16-bit modes: keep raw conv outputs (read once, by GroupNorm-apply) in fp16
    shared_prefix: bool = True   # CFG batching: inc / down1 convolutions computed once for both halves
    fused_sa: bool = True        # fused SelfAttention head / tail kernels (C = 64 / 128, 16-bit modes)
    fused_outc: bool = True      # 1x1 output conv inside the last SelfAttention tail
    vcat: bool = True            # Up blocks: residual concat recomputed in GroupNorm-apply, only the 16-bit copy stored
    simt_attention: bool = False  # 16-bit modes: fp32 CUDA-core attention core (comparator)

EMB_BLOCKS = ("down1", "down2", "down3", "up1", "up2", "up3")  # order of the concatenated emb_layer projection
TIME_DIM = 256


PDL_MAX_ROWS = 128  # UNet rows (2 x batch under CFG) up to which the step kernels are launched with PDL


def _pack_gemm_weight(w: torch.Tensor, dtype, device, split: bool):
    """Conv [Cout, Cin, 3, 3] -> [9 taps (dy, dx), Cout, Cin] / Linear [Cout, Cin] -> [1, Cout, Cin], K (= Cin) contiguous;
    for the split-tf32 engine a (hi, lo) pair of fp32 tensors."""
    packed = ops.pack_weights(w.detach().to(device=device, dtype=torch.float32).contiguous(), dtype)
    return ops.split_tf32(packed) if split else packed


def _wshape(w):
    return (w[0] if isinstance(w, tuple) else w).shape


class PackedWeights:
    """Kernel-layout copy of a UNet_conditional state_dict.  Derived data: rebuild after load_state_dict."""

    def __init__(self, sd, device, mode: str):
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        self.mode = mode
        self.act = MODES[mode]
        self.tf32 = mode == "fp32"  # split-tf32 tensor-core engine: GEMM weights are (hi, lo) pairs
        self.device = torch.device(device)
        self.c_in = sd["inc.double_conv.0.weight"].shape[1]
        self.c_out = sd["outc.weight"].shape[0]
        self.deep = "bot2.double_conv.0.weight" in sd
        f32 = lambda k: sd[k].detach().to(device=self.device, dtype=torch.float32).contiguous()  # noqa: E731
        self.t = {}
        host = lambda k: sd[k].detach().to(device="cpu", dtype=torch.float32).contiguous()  # noqa: E731
        for k, v in sd.items():
            if k == "inc.double_conv.0.weight":
                # direct fp32 conv, reference layout [64, c_in, 3, 3]; a HOST tensor: the <= 9 KB travel by value as a
                # kernel launch parameter (constant-bank FMA operands owned by each launch)
                self.t[k] = host(k)
            elif k.endswith(".weight") and v.dim() == 4 and v.shape[-1] == 3:
                self.t[k] = _pack_gemm_weight(v, self.act, self.device, self.tf32)
            elif k.endswith("in_proj_weight") or k.endswith("out_proj.weight") or ".ff_self.1.weight" in k or ".ff_self.3.weight" in k:
                self.t[k] = _pack_gemm_weight(v, self.act, self.device, self.tf32)
            elif ".emb_layer." in k or k == "label_emb.weight":
                continue
            elif k == "outc.weight":
                self.t[k] = f32(k).reshape(self.c_out, -1).contiguous()
            else:
                self.t[k] = f32(k)  # norm affines, biases
        # host copies for the fused output conv (launch parameters, like inc.double_conv.0)
        self.outc_host = (host("outc.weight").reshape(self.c_out, -1).contiguous(), host("outc.bias"))
        self.w_emb = torch.cat([f32(f"{b}.emb_layer.1.weight") for b in EMB_BLOCKS], 0).contiguous()
        self.b_emb = torch.cat([f32(f"{b}.emb_layer.1.bias") for b in EMB_BLOCKS], 0).contiguous()
        self.emb_off = {}
        off = 0
        for b in EMB_BLOCKS:
            n = sd[f"{b}.emb_layer.1.bias"].shape[0]
            self.emb_off[b] = (off, n)
            off += n
        self.emb_total = off
        self.label = f32("label_emb.weight") if "label_emb.weight" in sd else None
        self.num_classes = None if self.label is None else self.label.shape[0]
        # exactly the reference expression (diff_modules.py:169), evaluated on the CPU like the oracle
        self.inv_freq = (1.0 / (10000 ** (torch.arange(0, TIME_DIM, 2).float() / TIME_DIM))).to(self.device)

    def __getitem__(self, k):
        return self.t[k]


class UNetPlan:
    """Launch plan for `rows` batch rows at S x S.  Inputs live in plan-owned buffers:
    x_in fp32 NCHW [n_src, c_in, S, S] (row r reads sample r % n_src), t fp32 [rows] or the device step
    counter, y int64 [rows] (negative = unconditional).  Output: eps fp32 NCHW [rows, c_out, S, S]."""

    def __init__(self, weights: PackedWeights, *, n_src: int, rows: int, S: int, use_step: bool = False,
                 debug: bool = False, options: PlanOptions = PlanOptions()):
        if S < 16 or S & (S - 1):
            raise ValueError(f"img size {S}: the sgb200 kernels need a power-of-two size >= 16")
        if rows % n_src:
            raise ValueError("rows must be a multiple of n_src")
        self.W = weights
        self.dev = weights.device
        _cabi.require_b200(self.dev)
        self.opt = options
        self.tc = weights.mode in ("bf16", "f16")  # 16-bit operand copies next to the fp32 residual stream
        self.tf32 = weights.tf32                   # fp32 activations, split into (hi, lo) in front of every GEMM
        self.act = weights.act
        self.engine = SG_ENGINE_SIMT if weights.mode == "fp32_simt" else SG_ENGINE_TC
        self.rows, self.n_src, self.S = rows, n_src, S
        self.use_step = use_step
        self.raw16 = self.tc and options.raw16  # fp16 raw conv outputs in the 16-bit modes
        # CFG batching puts the conditional rows in [0, n) and the unconditional ones in [n, 2n); both halves see the
        # same x, and `inc` / the convolutions of `down1` take no embedding (:187-188, :110-113: emb is added AFTER the
        # convs), so that prefix is computed once for n rows and broadcast where the embedding comes in
        self.rows_p = n_src if (rows == 2 * n_src and options.shared_prefix) else rows
        self.debug = debug  # keep every buffer alive and expose per-block outputs in self.taps
        self._pool = {}
        self._lo = {}  # split-tf32 engine: data_ptr of an fp32 activation -> its lo part, until the consuming GEMM
        self.nbytes = 0
        self.ops = []
        self.n_launches = 0
        f32, dev = torch.float32, self.dev
        self.x_in = torch.zeros((n_src, weights.c_in, S, S), dtype=f32, device=dev)
        self.t = torch.zeros((rows,), dtype=f32, device=dev)
        self.y = torch.full((rows,), -1, dtype=torch.int64, device=dev)
        self.step = torch.zeros((1,), dtype=torch.int32, device=dev)
        # raised by GroupNorm-apply when an fp16 raw conv output left fp16's safe range (GroupNorm is scale invariant,
        # fp16 is not): the caller re-runs with PlanOptions(raw16=False).  Checked by the host after a run, never inside it.
        self.range_flag = torch.zeros((1,), dtype=torch.int32, device=dev) if self.raw16 else None
        self.temb = torch.empty((rows, TIME_DIM), dtype=f32, device=dev)
        self.emb = torch.empty((rows, weights.emb_total), dtype=f32, device=dev)
        self.eps = torch.empty((rows, weights.c_out, S, S), dtype=f32, device=dev)
        self.taps = {}  # debug only: block name -> fp32 NHWC output buffer
        self._build()

    # ------------------------------------------------------------------ buffers
    def _alloc(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        key = (n * torch.empty((), dtype=dtype).element_size(), dtype)
        lst = self._pool.get(key)
        if lst:
            return lst.pop().view(shape)
        t = torch.empty(shape, dtype=dtype, device=self.dev)
        self.nbytes += t.numel() * t.element_size()
        return t

    def _free(self, *ts):
        if self.debug:
            return
        for t in ts:
            if t is None:
                continue
            key = (t.numel() * t.element_size(), t.dtype)
            self._pool.setdefault(key, []).append(t)

    def _op(self, fn, *a, **kw):
        self.ops.append((fn, a, kw))
        self.n_launches += 1  # every entry point used by the plan is exactly one kernel launch

    def _gemm_in(self, t):
        """GEMM operand form of an activation buffer: the buffer itself, or -- split-tf32 engine -- the pair (t, lo): the
        fp32 tensor is its own high part and lo = tf32(t - trunc19(t)) was written next to it by the producing kernel
        (_lo_out) or, failing that, by one sg_split_tf32 pass.  Returns (operand, buffers to free after the consumer)."""
        if not self.tf32:
            return t, ()
        lo = self._lo.pop(t.data_ptr(), None)
        if lo is None:
            lo = self._alloc(t.shape, torch.float32)
            self._op(ops.split_tf32_lo, t, lo)
        return (t, lo), (lo,)

    def _lo_out(self, t):
        """split-tf32 engine: a lo buffer for the producer of fp32 tensor `t` to fill (as its fp32 `out_act`); the GEMM
        that consumes t picks it up in _gemm_in.  None in the other engines."""
        if not self.tf32:
            return None
        lo = self._alloc(t.shape, torch.float32)
        self._lo[t.data_ptr()] = lo
        return lo

    def _pair(self, shape, want_f32=True, want_act=True):
        """(fp32 buffer, GEMM-operand buffer) for one logical tensor; in fp32 mode they are the same buffer."""
        if not self.tc:
            b = self._alloc(shape, torch.float32)
            return b, b
        return (self._alloc(shape, torch.float32) if want_f32 else None,
                self._alloc(shape, self.act) if want_act else None)

    def _free_pair(self, p):
        if p[0] is p[1]:
            self._free(p[0])
        else:
            self._free(p[0], p[1])

    # ------------------------------------------------------------------ blocks
    def _conv(self, a_act, wname, rows, H, W):
        """3x3 conv (no bias) -> (raw fp32 [rows,H,W,Cout], GroupNorm partials)."""
        w = self.W[wname]
        cout = _wshape(w)[1]
        # 16-bit modes keep the raw conv output (it is only ever read by GroupNorm-apply) in fp16: half the
        # bytes of the conv epilogue and of the normalisation pass; the statistics come from the fp32 accumulators
        raw16 = self.raw16
        raw = self._alloc((rows, H, W, cout), torch.float16 if raw16 else torch.float32)
        P = ops.igemm_partials(self.engine, H, W, cout)
        part = self._alloc((rows, P, 2), torch.float32)
        a_op, tmp = self._gemm_in(a_act)
        self._op(ops.igemm, a_op, w, rows=rows, H=H, W=W, partials=part,
                 **({"out_act": raw} if raw16 else {"out_f32": raw}))
        self._free(*tmp)
        return raw, part

    def _double_conv(self, p, x, rows, H, W, *, residual=False, emb=None, want_f32=True, want_act=True, from_input=False,
                     out_rows=None):
        """DoubleConv (:75-93).  x = (fp32, act) pair of the input (or None when from_input).  Returns a pair.
        out_rows > rows: the last GroupNorm-apply (+emb) writes out_rows rows, row r from raw row r % rows."""
        W_ = self.W
        if from_input:
            raw1 = self._alloc((rows, H, W, 64), torch.float16 if self.raw16 else torch.float32)
            part1 = self._alloc((rows, ops.conv_in_partials(H), 2), torch.float32)
            self._op(ops.conv_in, self.x_in, W_[f"{p}.double_conv.0.weight"], raw1, part1)
        else:
            raw1, part1 = self._conv(x[1], f"{p}.double_conv.0.weight", rows, H, W)
        mid = self._pair(raw1.shape, want_f32=False, want_act=True)
        self._op(ops.gn_apply, raw1, part1, W_[f"{p}.double_conv.1.weight"], W_[f"{p}.double_conv.1.bias"], mode=1,
                 out_f32=None if self.tc else mid[0], out_act=mid[1] if self.tc else self._lo_out(mid[0]),
                 range_flag=self.range_flag)
        self._free(raw1, part1)
        raw2, part2 = self._conv(mid[1], f"{p}.double_conv.3.weight", rows, H, W)
        self._free_pair(mid)
        out = self._pair((out_rows or rows,) + tuple(raw2.shape[1:]), want_f32, want_act)
        if isinstance(residual, tuple):
            # Up's first DoubleConv: the residual is cat([skip, upsample(x)]), recomputed from (x, skip) in the kernel
            _, x_src, skip_src = residual
            self._op(ops.gn_apply_vcat, raw2, part2, W_[f"{p}.double_conv.4.weight"], W_[f"{p}.double_conv.4.bias"],
                     x_src, skip_src, out[1], range_flag=self.range_flag)
        else:
            self._op(ops.gn_apply, raw2, part2, W_[f"{p}.double_conv.4.weight"], W_[f"{p}.double_conv.4.bias"],
                     mode=2 if residual else 0, residual=x[0] if residual else None, emb=emb,
                     out_f32=out[0], out_act=out[1] if self.tc else (self._lo_out(out[0]) if want_act else None),
                     range_flag=self.range_flag)
        self._free(raw2, part2)
        return out

    def _emb_slice(self, block):
        off, n = self.W.emb_off[block]
        return self.emb[:, off:off + n]

    def _self_attention(self, p, x, rows, H, W, C, *, want_act, outc=None):
        """SelfAttention (:52-72) on the fp32 residual stream x [rows, H, W, C] (tokens are already row-major).
        outc = (w, b, eps): the block is the last one and the fused tail applies the 1x1 output conv itself; returns
        None when it did (eps is written, the block output is only materialised in debug mode)."""
        W_ = self.W
        L = H * W
        M = rows * L
        f32 = torch.float32
        tc_attn = self.tc and not self.opt.simt_attention  # tcgen05 kind::f16 attention core (16-bit q / k / v)
        fused = tc_attn and C in ops.FUSED_TOKEN_C and self.opt.fused_sa
        lin_dt = self.act if self.tc else f32  # dtype of the Linear layers' activations

        def linear(a, wname, bname, **kw):
            a_op, tmp = self._gemm_in(a)
            self._op(ops.igemm, a_op, W_[wname], rows=rows, H=H, W=W, bias=W_[bname], **kw)
            self._free(*tmp)

        # tcgen05 attention consumes 16-bit q/k/v; the split-tf32 and the CUDA-core cores read fp32
        qkv = self._alloc((M, 3 * C), self.act if tc_attn else f32)
        if fused:
            # LayerNorm + in_proj in one pass over x (sg_ln_inproj)
            self._op(ops.ln_inproj, x, W_[f"{p}.ln.weight"], W_[f"{p}.ln.bias"], W_[f"{p}.mha.in_proj_weight"],
                     W_[f"{p}.mha.in_proj_bias"], qkv)
        else:
            ln1 = self._alloc((M, C), lin_dt)
            self._op(ops.layernorm, x, W_[f"{p}.ln.weight"], W_[f"{p}.ln.bias"], ln1)
            linear(ln1, f"{p}.mha.in_proj_weight", f"{p}.mha.in_proj_bias",
                   **({"out_act": qkv} if tc_attn else {"out_f32": qkv}))
            self._free(ln1)
        att = self._alloc((M, C), lin_dt)
        if self.tf32 and L >= 128:
            # split-TF32 core: (hi, lo) of q | k and of V transposed per (row, head), one preparation pass over qkv
            qk_hi, qk_lo = self._alloc((M, 2 * C), f32), self._alloc((M, 2 * C), f32)
            vt_hi, vt_lo = self._alloc((rows * C, L), f32), self._alloc((rows * C, L), f32)
            self._op(ops.attn_prep_tf32, qkv, qk_hi, qk_lo, vt_hi, vt_lo, rows=rows, L=L, C=C)
            self._op(ops.attention_tf32, qk_hi, qk_lo, vt_hi, vt_lo, att, rows=rows, L=L, C=C)
            self._free(qk_hi, qk_lo, vt_hi, vt_lo)
        else:
            # (fp32 engines at L < 128: a handful of tokens per row, the CUDA-core kernel)
            self._op(ops.attention, qkv, att, rows=rows, L=L, C=C, engine=SG_ENGINE_TC if tc_attn else SG_ENGINE_SIMT)
        self._free(qkv)
        if fused and not want_act:
            # out_proj + residual + LayerNorm + FFN + residual in one pass (sg_attn_tail)
            fuse_outc = outc is not None and C == 64 and self.opt.fused_outc
            out = self._pair((rows, H, W, C), True, False) if (not fuse_outc or self.debug) else (None, None)
            self._op(ops.attn_tail, att, x, W_[f"{p}.mha.out_proj.weight"], W_[f"{p}.mha.out_proj.bias"],
                     W_[f"{p}.ff_self.0.weight"], W_[f"{p}.ff_self.0.bias"], W_[f"{p}.ff_self.1.weight"],
                     W_[f"{p}.ff_self.1.bias"], W_[f"{p}.ff_self.3.weight"], W_[f"{p}.ff_self.3.bias"], out[0],
                     **({"outc": (W_.outc_host[0], W_.outc_host[1], outc[2])} if fuse_outc else {}))
            self._free(att)
            if fuse_outc:
                if self.debug:
                    self.taps_last = out[0]
                return None
            return out
        a = self._alloc((rows, H, W, C), f32)
        linear(att, f"{p}.mha.out_proj.weight", f"{p}.mha.out_proj.bias", residual=x, out_f32=a)
        self._free(att)
        ln2 = self._alloc((M, C), lin_dt)
        self._op(ops.layernorm, a, W_[f"{p}.ff_self.0.weight"], W_[f"{p}.ff_self.0.bias"], ln2)
        f1 = self._alloc((M, C), lin_dt)
        linear(ln2, f"{p}.ff_self.1.weight", f"{p}.ff_self.1.bias", gelu=True,
               **({"out_act": f1} if self.tc else {"out_f32": f1}))
        self._free(ln2)
        out = self._pair((rows, H, W, C), True, want_act)
        linear(f1, f"{p}.ff_self.3.weight", f"{p}.ff_self.3.bias", residual=a,
               out_f32=out[0], out_act=out[1] if self.tc else None)
        self._free(f1, a)
        return out

    def _down(self, p, x, rows, H, W, C, *, out_rows=None):
        h, w = H // 2, W // 2
        pooled = self._pair((rows, h, w, C))
        self._op(ops.maxpool2, x[0], out_f32=pooled[0], out_act=pooled[1] if self.tc else self._lo_out(pooled[0]))
        d1 = self._double_conv(f"{p}.maxpool_conv.1", pooled, rows, h, w, residual=True, want_f32=False)
        self._free_pair(pooled)
        d2 = self._double_conv(f"{p}.maxpool_conv.2", d1, rows, h, w, emb=self._emb_slice(p), want_act=False,
                               out_rows=out_rows)
        self._free_pair(d1)
        return d2

    def _up(self, p, x, skip, rows, h, w):
        H, W = 2 * h, 2 * w
        ct = x[0].shape[-1] + skip[0].shape[-1]
        # tensor-core modes with fp16 raw tensors: only the 16-bit operand copy of the concatenation is materialised;
        # the residual DoubleConv recomputes its fp32 residual from (x, skip) (sg_gn_apply_vcat)
        vcat = self.raw16 and x[0].shape[-1] == skip[0].shape[-1] and self.opt.vcat
        cat = self._pair((rows, H, W, ct), want_f32=not vcat)
        self._op(ops.upsample_cat, x[0], skip[0], out_f32=cat[0], out_act=cat[1] if self.tc else self._lo_out(cat[0]))
        u1 = self._double_conv(f"{p}.conv.0", cat, rows, H, W, residual=("vcat", x[0], skip[0]) if vcat else True,
                               want_f32=False)
        self._free_pair(cat)
        u2 = self._double_conv(f"{p}.conv.1", u1, rows, H, W, emb=self._emb_slice(p), want_act=False)
        self._free_pair(u1)
        return u2

    # ------------------------------------------------------------------ whole network (:175-196)
    def _build(self):
        W_, rows, S = self.W, self.rows, self.S
        keep = self.taps
        self._op(ops.time_embed, None if self.use_step else self.t, self.step if self.use_step else None,
                 self.y if W_.label is not None else None,
                 W_.inv_freq, W_.label, W_.w_emb, W_.b_emb, self.temb, self.emb)
        rp = self.rows_p  # rows of the label-independent prefix (n when the CFG halves are batched, else rows)
        x1 = self._double_conv("inc", None, rp, S, S, from_input=True, want_act=False)
        keep["inc"] = x1[0]
        d = self._down("down1", x1, rp, S, S, 64, out_rows=rows)
        keep["down1"] = d[0]
        x2 = self._self_attention("sa1", d[0], rows, S // 2, S // 2, 128, want_act=False)
        self._free_pair(d)
        keep["sa1"] = x2[0]
        d = self._down("down2", x2, rows, S // 2, S // 2, 128)
        keep["down2"] = d[0]
        x3 = self._self_attention("sa2", d[0], rows, S // 4, S // 4, 256, want_act=False)
        self._free_pair(d)
        keep["sa2"] = x3[0]
        d = self._down("down3", x3, rows, S // 4, S // 4, 256)
        keep["down3"] = d[0]
        x4 = self._self_attention("sa3", d[0], rows, S // 8, S // 8, 256, want_act=True)
        self._free_pair(d)
        keep["sa3"] = x4[0]
        s8 = S // 8
        b = self._double_conv("bot1", x4, rows, s8, s8, want_f32=self.debug)
        keep["bot1"] = b[0]
        self._free_pair(x4)
        if W_.deep:
            b2 = self._double_conv("bot2", b, rows, s8, s8, want_f32=self.debug)
            keep["bot2"] = b2[0]
            self._free_pair(b)
            b = b2
        b3 = self._double_conv("bot3", b, rows, s8, s8, want_act=False)
        keep["bot3"] = b3[0]
        self._free_pair(b)
        u = self._up("up1", b3, x3, rows, s8, s8)
        self._free_pair(b3)
        self._free_pair(x3)
        keep["up1"] = u[0]
        a = self._self_attention("sa4", u[0], rows, S // 4, S // 4, 128, want_act=False)
        self._free_pair(u)
        keep["sa4"] = a[0]
        u = self._up("up2", a, x2, rows, S // 4, S // 4)
        self._free_pair(a)
        self._free_pair(x2)
        keep["up2"] = u[0]
        a = self._self_attention("sa5", u[0], rows, S // 2, S // 2, 64, want_act=False)
        self._free_pair(u)
        keep["sa5"] = a[0]
        u = self._up("up3", a, x1, rows, S // 2, S // 2)
        self._free_pair(a)
        self._free_pair(x1)
        keep["up3"] = u[0]
        a = self._self_attention("sa6", u[0], rows, S, S, 64, want_act=False,
                                 outc=(W_["outc.weight"], W_["outc.bias"], self.eps))
        self._free_pair(u)
        if a is None:  # the fused tail applied outc itself
            if self.debug:
                keep["sa6"] = self.taps_last
        else:
            keep["sa6"] = a[0]
            self._op(ops.conv_out, a[0].view(rows, S * S, 64), W_["outc.weight"], W_["outc.bias"], self.eps)
            self._free_pair(a)
        assert not self._lo, "a lo operand was produced for a tensor no GEMM consumed"
        self._pool.clear()

    def run(self):
        """Issue the whole forward on the current stream of the plan's device (graph-capturable)."""
        with torch.cuda.device(self.dev):
            # programmatic dependent launch pays off while the step is launch-latency bound (measured break-even: batch 64)
            ops.set_pdl(2 if self.rows <= PDL_MAX_ROWS else 0)
            for fn, a, kw in self.ops:
                fn(*a, **kw)

    def range_overflow(self) -> bool:
        """True when a GroupNorm input of a run since the last reset left fp16's safe range (host sync; the results of
        that run are then not trustworthy and the caller rebuilds the plan with PlanOptions(raw16=False))."""
        return self.range_flag is not None and bool(self.range_flag.item())

    def reset_range_flag(self):
        if self.range_flag is not None:
            self.range_flag.zero_()
