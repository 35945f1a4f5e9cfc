"""DiffusionVAE decode tail on the sgb200 kernels (/root/reference/src/diff_modules.py:702-706):
clamp(-1,1) -> VQEmbeddingEMA (nearest codeword) -> Decoder -> (x+1)/2*255 -> uint8, the step that turns the sampled
[n,4,S,S] latents into the [n,1,4S,4S] spectrograms the shipped generator writes.

Launch sequence per chunk of samples (all through the C ABI, no torch-op fallback):
  sg_vq_quantize  -> sg_dec_in_proj (1x1, 4 -> 512)
  -> sg_igemm taps=1 (+bias +residual, ReLU)        residual_conv_1   (:342-344)
  -> sg_igemm taps=9 (+bias +residual, ReLU)        residual_conv_2   (:346-348)
  -> 2 x sg_igemm taps=1 (512 -> 2 x 512, +bias)    strided_t_conv_1 as a Linear per input pixel, one launch per
                                                    output row parity a; its output is left un-shuffled
  -> sg_tconv2_u8                                   strided_t_conv_2 + image tail
"""
from __future__ import annotations

import torch

from . import _cabi, ops
from ._cabi import SG_ENGINE_SIMT, SG_ENGINE_TC
from .engine import MODES, _pack_gemm_weight

HIDDEN, LATENT = 512, 4


class VqaeDecoder:
    """Packed decoder + codebook of a reference VQAE state_dict (keys `codebook.embedding`, `decoder.*`; encoder and
    EMA buffers are ignored).  Derived data: rebuild after the state_dict changes."""

    def __init__(self, state_dict, device, mode="bf16"):
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _cabi.require_b200(self.device)
        self.mode, self.act = mode, MODES[mode]
        self.tc = mode in ("bf16", "f16")  # 16-bit operand copies
        self.tf32 = mode == "fp32"         # fp32-accurate tensor-core engine: (hi, lo) split operands
        self.engine = SG_ENGINE_SIMT if mode == "fp32_simt" else SG_ENGINE_TC
        wsplit = (lambda t: ops.split_tf32(t)) if self.tf32 else (lambda t: t)
        sd = state_dict
        f32 = lambda k: sd[k].detach().to(device=self.device, dtype=torch.float32).contiguous()  # noqa: E731
        for k, shape in (("codebook.embedding", (None, LATENT)), ("decoder.in_proj.weight", (HIDDEN, LATENT, 1, 1)),
                         ("decoder.residual_conv_1.weight", (HIDDEN, HIDDEN, 1, 1)),
                         ("decoder.residual_conv_2.weight", (HIDDEN, HIDDEN, 3, 3)),
                         ("decoder.strided_t_conv_1.weight", (HIDDEN, HIDDEN, 2, 2)),
                         ("decoder.strided_t_conv_2.weight", (HIDDEN, 1, 2, 2))):
            if k not in sd:
                raise KeyError(f"VQAE state_dict lacks {k}")
            got = tuple(sd[k].shape)
            if any(s is not None and s != g for s, g in zip(shape, got)) or len(got) != len(shape):
                raise ValueError(f"{k}: shape {got}, expected {shape}")
        self.codebook = f32("codebook.embedding")
        self.w_in = f32("decoder.in_proj.weight").reshape(HIDDEN, LATENT).contiguous()
        self.b_in = f32("decoder.in_proj.bias")
        self.w_r1 = wsplit(f32("decoder.residual_conv_1.weight").reshape(1, HIDDEN, HIDDEN).contiguous().to(self.act))
        self.b_r1 = f32("decoder.residual_conv_1.bias")
        self.w_r2 = _pack_gemm_weight(sd["decoder.residual_conv_2.weight"], self.act, self.device, self.tf32)
        self.b_r2 = f32("decoder.residual_conv_2.bias")
        # ConvTranspose2d weight [ci, co, a, b] -> for each a: Linear [(b, co), ci]
        wt = f32("decoder.strided_t_conv_1.weight")
        self.w_t1 = [wsplit(wt[:, :, a, :].permute(2, 1, 0).reshape(1, 2 * HIDDEN, HIDDEN).contiguous().to(self.act))
                     for a in (0, 1)]
        self.b_t1 = f32("decoder.strided_t_conv_1.bias").repeat(2).contiguous()
        self.w_t2 = f32("decoder.strided_t_conv_2.weight")
        self.b_t2 = f32("decoder.strided_t_conv_2.bias")
        self._bufs = {}
        self.gpu_launches = 0

    def _buffers(self, n, S):
        key = (n, S)
        b = self._bufs.get(key)
        if b is None:
            dev, f32 = self.device, torch.float32
            M = n * S * S
            b = {"q": torch.empty((n, LATENT, S, S), dtype=f32, device=dev),
                 "idx": torch.empty((n * LATENT * S * S // 4,), dtype=torch.int32, device=dev),
                 "h0f": torch.empty((n, S, S, HIDDEN), dtype=f32, device=dev),
                 "h1f": torch.empty((n, S, S, HIDDEN), dtype=f32, device=dev),
                 "t": torch.empty((2, M, 2 * HIDDEN), dtype=self.act, device=dev)}
            if self.tc:
                for k in ("h0a", "h1a", "h2a"):
                    b[k] = torch.empty((n, S, S, HIDDEN), dtype=self.act, device=dev)
            else:
                b["h0a"], b["h1a"] = b["h0f"], b["h1f"]
                b["h2a"] = torch.empty((n, S, S, HIDDEN), dtype=f32, device=dev)
            if self.tf32:
                b["hi"], b["lo"] = (torch.empty((n, S, S, HIDDEN), dtype=f32, device=dev) for _ in range(2))
            self._bufs = {key: b}  # keep one geometry resident
        return b

    @torch.no_grad()
    def decode(self, x, *, micro_batch=64, return_float=False, return_indices=False, return_quantized=False):
        """x: fp32 [n, 4, S, S] sampler state -> uint8 [n, 1, 4S, 4S] (or the fp32 decoder output).  x is not modified.
        return_indices / return_quantized append the codeword indices / the quantised fp32 latents [n, 4, S, S] (the
        `x + (q - x)` value of :313) to the result."""
        if x.device != self.device or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != LATENT:
            raise ValueError("decode: x must be an fp32 [n, 4, S, S] tensor on the decoder's device")
        n, _, S, _ = x.shape
        x = x.contiguous()
        out = torch.empty((n, 1, 4 * S, 4 * S), dtype=torch.float32 if return_float else torch.uint8, device=self.device)
        all_idx = torch.empty((n * LATENT * S * S // 4,), dtype=torch.int32, device=self.device) if return_indices else None
        all_q = torch.empty((n, LATENT, S, S), dtype=torch.float32, device=self.device) if return_quantized else None
        self.gpu_launches = 0
        for lo in range(0, n, micro_batch):
            hi = min(n, lo + micro_batch)
            nb = hi - lo
            b = self._buffers(nb, S)
            ops.vq_quantize(x[lo:hi], self.codebook, b["q"], b["idx"], clamp=True)
            tc = self.tc
            ops.dec_in_proj(b["q"], self.w_in, self.b_in, out_f32=b["h0f"], out_act=b["h0a"] if tc else None)
            kw = dict(rows=nb, H=S, W=S)
            # split-tf32 engine: every GEMM operand goes through one sg_split_tf32 pass into the (hi, lo) scratch pair
            opnd = (lambda t: ops.split_tf32(t, b["hi"], b["lo"])) if self.tf32 else (lambda t: t)
            ops.igemm(opnd(b["h0a"]), self.w_r1, bias=self.b_r1, residual=b["h0f"], relu_post=True, out_f32=b["h1f"],
                      out_act=b["h1a"] if tc else None, **kw)
            ops.igemm(opnd(b["h1a"]), self.w_r2, bias=self.b_r2, residual=b["h1f"], relu_post=True,
                      **({"out_act": b["h2a"]} if tc else {"out_f32": b["h2a"]}), **kw)
            h2 = opnd(b["h2a"])
            for a in (0, 1):
                ops.igemm(h2, self.w_t1[a], bias=self.b_t1, **({"out_act": b["t"][a]} if tc else {"out_f32": b["t"][a]}),
                          **kw)
            ops.tconv2_u8(b["t"], self.w_t2, self.b_t2, n=nb, S=S, **({"out_f32": out[lo:hi]} if return_float else
                                                                      {"out_u8": out[lo:hi]}))
            self.gpu_launches += 10 if self.tf32 else 7
            if return_indices:
                g = nb * LATENT * S * S // 4
                all_idx[lo * LATENT * S * S // 4: lo * LATENT * S * S // 4 + g].copy_(b["idx"][:g])
            if return_quantized:
                all_q[lo:hi].copy_(b["q"])
        res = (out,) + ((all_idx,) if return_indices else ()) + ((all_q,) if return_quantized else ())
        return res if len(res) > 1 else out
