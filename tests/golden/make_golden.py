"""Golden-vector generator: runs the UNMODIFIED reference (imported from /root/reference/src)
on CPU and stores its outputs as small fixtures next to this script.

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Weights and inputs are NOT stored: they are rebuilt from seeds by oracle/weights.py and
`golden_inputs` below with the CPU generator (same torch build on both boxes).

Import recipe: the reference imports matplotlib at module top
(diff_modules.py:4, diff_utils.py:9) only for plotting, and matplotlib is absent here, so two
empty modules are registered before the import.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.weights import make_state_dict  # noqa: E402

REF_SRC = "/root/reference/src"

# (tag, S, c, n, t-values) of the single-step eps fixtures
EPS_CASES = [
    ("r16", 16, 4, 2, (999, 500, 1)),
    ("r64", 64, 4, 2, (999, 20)),
    ("p32c1", 32, 1, 1, (300,)),  # pixel-space variant (c_in = c_out = 1)
]
TRAJ_CASES = [("r16", 16, 4, 4, 50), ("r64", 64, 4, 2, 50)]
NUM_CLASSES = 27
WEIGHT_SEED = 1234


def golden_inputs(s, c, n, seed=123):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.randn((n, c, s, s), generator=g)
    y = torch.randint(0, NUM_CLASSES, (n,), generator=g)
    return x, y


def import_reference():
    os.environ.setdefault("WANDB_MODE", "disabled")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.path.insert(0, REF_SRC)
    import diff_modules  # noqa
    import diff_utils  # noqa

    diff_modules.progress_bar = lambda it, **kw: it
    return diff_modules, diff_utils


def main():
    torch.set_num_threads(os.cpu_count())
    dm, du = import_reference()
    out = {}

    # ---- state_dict schema of reference-constructed models -------------------------------
    schema = {}
    for tag, kw in {
        "c4_k27": dict(c_in=4, c_out=4, num_classes=27),
        "c1_k10": dict(c_in=1, c_out=1, num_classes=10),
        "c4_k27_shallow": dict(c_in=4, c_out=4, num_classes=27, remove_deep_conv=True),
    }.items():
        m = dm.UNet_conditional(**kw)
        schema[tag] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    with open(os.path.join(HERE, "state_dict_schema.json"), "w") as f:
        json.dump(schema, f)

    # ---- schedule known answers ----------------------------------------------------------
    d = dm.Diffusion(noise_steps=1000, img_size=16, num_classes=NUM_CLASSES, c_in=4, c_out=4, device="cpu")
    out["sched_beta"] = d.beta.numpy()
    out["sched_alpha_hat"] = d.alpha_hat.numpy()

    # ---- pos_encoding --------------------------------------------------------------------
    m = dm.UNet_conditional(4, 4, num_classes=NUM_CLASSES)
    tt = torch.tensor([0, 1, 2, 20, 500, 998, 999]).long()
    out["posenc_t"] = tt.numpy()
    out["posenc"] = m.pos_encoding(tt.unsqueeze(-1), 256).numpy()

    # ---- single-step eps (cond + uncond), per-block taps at r16 --------------------------
    for tag, s, c, n, ts in EPS_CASES:
        sd = make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES)
        m = dm.UNet_conditional(c, c, num_classes=NUM_CLASSES)
        m.load_state_dict(sd, strict=True)
        m.eval()
        x, y = golden_inputs(s, c, n)
        for tv in ts:
            t = (torch.ones(n) * tv).long()
            with torch.inference_mode():
                out[f"eps_{tag}_t{tv}_cond"] = m(x, t, y).numpy()
                out[f"eps_{tag}_t{tv}_uncond"] = m(x, t, None).numpy()
        if tag == "r16":
            taps = {}
            hooks = []
            for name in ["inc", "down1", "sa1", "down2", "sa2", "down3", "sa3", "bot1", "bot2", "bot3",
                         "up1", "sa4", "up2", "sa5", "up3", "sa6"]:
                hooks.append(getattr(m, name).register_forward_hook(
                    lambda mod, inp, o, name=name: taps.__setitem__(name, o.detach().clone())))
            t = (torch.ones(n) * 500).long()
            with torch.inference_mode():
                m(x[:1], t[:1], y[:1])
            for h in hooks:
                h.remove()
            for k, v in taps.items():
                out[f"tap_r16_{k}"] = v.numpy()

    # ---- remove_deep_conv variant ---------------------------------------------------------
    sd = make_state_dict(WEIGHT_SEED, 4, 4, NUM_CLASSES, remove_deep_conv=True)
    m = dm.UNet_conditional(4, 4, num_classes=NUM_CLASSES, remove_deep_conv=True)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x, y = golden_inputs(16, 4, 2)
    with torch.inference_mode():
        out["eps_r16shallow_t500_cond"] = m(x, (torch.ones(2) * 500).long(), y).numpy()

    # ---- trajectories: reference Diffusion.sample under set_seed --------------------------
    for tag, s, c, n, T in TRAJ_CASES:
        sd = make_state_dict(WEIGHT_SEED, c, c, NUM_CLASSES)
        d = dm.Diffusion(noise_steps=T, img_size=s, num_classes=NUM_CLASSES, c_in=c, c_out=c, device="cpu")
        d.model.load_state_dict(sd, strict=True)
        _, y = golden_inputs(s, c, n)
        # capture the float state before quantisation: the only clamp in the path is :440
        captured = {}
        orig_clamp = torch.Tensor.clamp

        def spy(self, *a, **k):
            captured["x"] = self.detach().clone()
            return orig_clamp(self, *a, **k)

        torch.Tensor.clamp = spy
        try:
            du.set_seed(7)
            u8 = d.sample(False, y, cfg_scale=3)
        finally:
            torch.Tensor.clamp = orig_clamp
        out[f"traj_{tag}_T{T}_u8"] = u8.numpy()
        out[f"traj_{tag}_T{T}_xfloat"] = captured["x"].numpy()
        out[f"traj_{tag}_T{T}_labels"] = y.numpy()
    # cfg_scale = 0 branch (:426): single conditional forward per step
    sd = make_state_dict(WEIGHT_SEED, 4, 4, NUM_CLASSES)
    d = dm.Diffusion(noise_steps=12, img_size=16, num_classes=NUM_CLASSES, c_in=4, c_out=4, device="cpu")
    d.model.load_state_dict(sd, strict=True)
    _, y = golden_inputs(16, 4, 2)
    du.set_seed(7)
    out["traj_r16_T12_cfg0_u8"] = d.sample(False, y, cfg_scale=0).numpy()

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    sizes = {k: int(v.nbytes) for k, v in out.items()}
    print(json.dumps({"n_arrays": len(out), "bytes": sum(sizes.values())}))


if __name__ == "__main__":
    main()
