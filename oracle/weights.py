"""Deterministic synthetic weights for UNet_conditional -- TEST INFRASTRUCTURE ONLY.

The reference ships no checkpoints, and a 93 MB state_dict cannot be
committed, so both the golden-vector generator (run in the build container with the
reference imported) and the GPU-side parity tests (run on a box without the reference)
rebuild the SAME weights from a seed with the CPU generator of the same torch build.

Schema follows /root/reference/src/diff_modules.py:75-217 (DoubleConv :82-86, Down
:99-108, Up :120-129, SelfAttention :56-63, UNet :144-166, label_emb :208): 183 tensors
at the default depth.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

TIME_DIM = 256  # hard-wired emb_dim of Down/Up (diff_modules.py:97,117) == time_dim (:140)


def _double_conv(prefix, cin, cout, cmid=None):
    cmid = cmid or cout
    return [
        (f"{prefix}.double_conv.0.weight", (cmid, cin, 3, 3)),
        (f"{prefix}.double_conv.1.weight", (cmid,)),
        (f"{prefix}.double_conv.1.bias", (cmid,)),
        (f"{prefix}.double_conv.3.weight", (cout, cmid, 3, 3)),
        (f"{prefix}.double_conv.4.weight", (cout,)),
        (f"{prefix}.double_conv.4.bias", (cout,)),
    ]


def _down(prefix, cin, cout):
    return (
        _double_conv(f"{prefix}.maxpool_conv.1", cin, cin)
        + _double_conv(f"{prefix}.maxpool_conv.2", cin, cout)
        + [(f"{prefix}.emb_layer.1.weight", (cout, TIME_DIM)), (f"{prefix}.emb_layer.1.bias", (cout,))]
    )


def _up(prefix, cin, cout):
    return (
        _double_conv(f"{prefix}.conv.0", cin, cin)
        + _double_conv(f"{prefix}.conv.1", cin, cout, cin // 2)
        + [(f"{prefix}.emb_layer.1.weight", (cout, TIME_DIM)), (f"{prefix}.emb_layer.1.bias", (cout,))]
    )


def _sa(prefix, c):
    return [
        (f"{prefix}.mha.in_proj_weight", (3 * c, c)),
        (f"{prefix}.mha.in_proj_bias", (3 * c,)),
        (f"{prefix}.mha.out_proj.weight", (c, c)),
        (f"{prefix}.mha.out_proj.bias", (c,)),
        (f"{prefix}.ln.weight", (c,)),
        (f"{prefix}.ln.bias", (c,)),
        (f"{prefix}.ff_self.0.weight", (c,)),
        (f"{prefix}.ff_self.0.bias", (c,)),
        (f"{prefix}.ff_self.1.weight", (c, c)),
        (f"{prefix}.ff_self.1.bias", (c,)),
        (f"{prefix}.ff_self.3.weight", (c, c)),
        (f"{prefix}.ff_self.3.bias", (c,)),
    ]


def state_dict_schema(c_in=4, c_out=4, num_classes=27, remove_deep_conv=False):
    """Ordered (name, shape) list in the reference's registration order."""
    s = []
    s += _double_conv("inc", c_in, 64)
    s += _down("down1", 64, 128) + _sa("sa1", 128)
    s += _down("down2", 128, 256) + _sa("sa2", 256)
    s += _down("down3", 256, 256) + _sa("sa3", 256)
    if remove_deep_conv:
        s += _double_conv("bot1", 256, 256) + _double_conv("bot3", 256, 256)
    else:
        s += _double_conv("bot1", 256, 512) + _double_conv("bot2", 512, 512) + _double_conv("bot3", 512, 256)
    s += _up("up1", 512, 128) + _sa("sa4", 128)
    s += _up("up2", 256, 64) + _sa("sa5", 64)
    s += _up("up3", 128, 64) + _sa("sa6", 64)
    s += [("outc.weight", (c_out, 64, 1, 1)), ("outc.bias", (c_out,))]
    if num_classes is not None:
        s += [("label_emb.weight", (num_classes, TIME_DIM))]
    return s


def make_state_dict(seed=0, c_in=4, c_out=4, num_classes=27, remove_deep_conv=False, init="perturbed"):
    """Synthetic fp32 weights, a pure function of (seed, schema, init).

    init="default":   same distributions as torch's default nn init (uniform(-1/sqrt(fan_in), ..)
                      for conv/linear, xavier for in_proj, unit affine, N(0,1) label table).
    init="perturbed": as above, but norm affines and the biases torch zero-initialises are
                      randomised too, so an implementation that drops one of them fails parity.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    for name, shape in state_dict_schema(c_in, c_out, num_classes, remove_deep_conv):
        leaf = name.rsplit(".", 1)[-1]
        is_norm = (
            (".double_conv.1." in name) or (".double_conv.4." in name) or (".ln." in name) or (".ff_self.0." in name)
        )
        if name == "label_emb.weight":
            w = torch.randn(shape, generator=g, dtype=torch.float32)
        elif is_norm:
            if init == "perturbed":
                r = torch.randn(shape, generator=g, dtype=torch.float32)
                w = 1.0 + 0.1 * r if leaf == "weight" else 0.1 * r
            else:
                w = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        elif name.endswith("in_proj_weight"):
            c = shape[1]
            w = uniform(shape, math.sqrt(6.0 / (c + 3 * c)))
        elif name.endswith("in_proj_bias") or name.endswith("out_proj.bias"):
            w = uniform(shape, 0.05) if init == "perturbed" else torch.zeros(shape)
        elif leaf == "weight":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = uniform(shape, 1.0 / math.sqrt(fan_in))
        else:  # biases of Linear / outc: fan_in of the matching weight
            wshape = dict(state_dict_schema(c_in, c_out, num_classes, remove_deep_conv))[name[: -len("bias")] + "weight"]
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            w = uniform(shape, 1.0 / math.sqrt(fan_in))
        sd[name] = w.contiguous()
    return sd
