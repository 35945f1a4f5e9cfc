"""CPU restatement of the DiffusionVAE decode tail -- TEST INFRASTRUCTURE ONLY.

Restates, functionally and state_dict-driven,
  * VQEmbeddingEMA.forward, eval path   /root/reference/src/diff_modules.py:290-318  (nearest codeword; note that
    `x.reshape(-1, D)` at :292 groups D = 4 consecutive elements of the contiguous NCHW tensor, i.e. 4 neighbouring
    W positions of one channel -- not the 4 channels of a pixel)
  * Decoder.forward                     :322-352  (1x1 in_proj, two residual convs with ReLU, two ConvTranspose2d k2 s2)
  * the tail of DiffusionVAE.sample     :702-706  (clamp(-1,1) -> codebook -> decoder -> (x+1)/2*255 -> uint8, NO clamp
    before the cast)
with the same torch CPU primitives the reference calls (torch.cdist, F.conv2d, F.conv_transpose2d), so it is pinned
bit-for-bit against the reference modules by tests/golden/make_golden_vae.py -> golden_vae.npz.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

HIDDEN = 512      # DiffusionVAE.__init__ :610
LATENT = 4        # :611
N_CODES = 512     # :612


def vqae_schema():
    """(name, shape) of the VQAE tensors the decode tail reads (decoder + codebook; :266-270, :326-334)."""
    return [
        ("codebook.embedding", (N_CODES, LATENT)),
        ("decoder.in_proj.weight", (HIDDEN, LATENT, 1, 1)),
        ("decoder.in_proj.bias", (HIDDEN,)),
        ("decoder.residual_conv_1.weight", (HIDDEN, HIDDEN, 1, 1)),
        ("decoder.residual_conv_1.bias", (HIDDEN,)),
        ("decoder.residual_conv_2.weight", (HIDDEN, HIDDEN, 3, 3)),
        ("decoder.residual_conv_2.bias", (HIDDEN,)),
        ("decoder.strided_t_conv_1.weight", (HIDDEN, HIDDEN, 2, 2)),  # ConvTranspose2d: [in, out, kH, kW]
        ("decoder.strided_t_conv_1.bias", (HIDDEN,)),
        ("decoder.strided_t_conv_2.weight", (HIDDEN, 1, 2, 2)),
        ("decoder.strided_t_conv_2.bias", (1,)),
    ]


def make_vqae_state_dict(seed=0):
    """Synthetic fp32 VQAE weights (a pure function of the seed, CPU generator).  The codebook spans [-1, 1] (the
    range of the clamped latents) instead of torch's +-1/512 init so that all codewords are in play, and the decoder
    weights are scaled so that the output image spans a good part of [-1, 1] -- and leaves it here and there, which
    exercises the un-clamped uint8 cast."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()
    for name, shape in vqae_schema():
        if name == "codebook.embedding":
            w = torch.rand(shape, generator=g) * 2 - 1
        elif name.endswith(".weight"):
            if "strided_t_conv" in name:
                fan_in = shape[0]  # every output pixel of a k2 s2 transposed conv sees exactly one tap
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            w = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
        else:
            w = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        sd[name] = w.float().contiguous()
    return sd


def vq_quantize(x, embedding):
    """VQEmbeddingEMA.forward in eval mode (:290-318): returns (quantized, indices).  `quantized` is evaluated as the
    reference's straight-through expression x + (q - x) (:313), which is not always bit-identical to q."""
    D = embedding.shape[1]
    x_flat = x.detach().reshape(-1, D)
    distances = (-torch.cdist(x_flat, embedding, p=2)) ** 2
    indices = torch.argmin(distances.float(), dim=-1)
    q = F.embedding(indices, embedding).view_as(x)
    return x + (q - x), indices


def decoder_forward(sd, z):
    """Decoder.forward (:338-352) on NCHW z [n, 4, S, S] -> [n, 1, 4S, 4S]."""
    p = "decoder."
    x = F.conv2d(z, sd[p + "in_proj.weight"], sd[p + "in_proj.bias"])
    y = F.conv2d(x, sd[p + "residual_conv_1.weight"], sd[p + "residual_conv_1.bias"])
    x = F.relu(y + x)
    y = F.conv2d(x, sd[p + "residual_conv_2.weight"], sd[p + "residual_conv_2.bias"], padding=1)
    y = F.relu(y + x)
    y = F.conv_transpose2d(y, sd[p + "strided_t_conv_1.weight"], sd[p + "strided_t_conv_1.bias"], stride=2)
    y = F.conv_transpose2d(y, sd[p + "strided_t_conv_2.weight"], sd[p + "strided_t_conv_2.bias"], stride=2)
    return y


def image_to_uint8(y):
    """(:704-705): (x + 1) / 2 * 255 -> truncating cast WITHOUT a clamp (the cast of an out-of-range float is
    whatever torch's CPU kernel does; the oracle calls that very kernel)."""
    return (((y + 1) / 2) * 255).type(torch.uint8)


def decode_tail(x, sd, return_all=False):
    """The tail of DiffusionVAE.sample (:702-706) on the float sampler state x [n, 4, S, S]."""
    xc = x.clamp(-1, 1)
    q, idx = vq_quantize(xc, sd["codebook.embedding"])
    y = decoder_forward(sd, q)
    u8 = image_to_uint8(y)
    if return_all:
        return u8, y, q, idx
    return u8
