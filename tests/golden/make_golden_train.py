"""Golden vectors of the forward-only training helpers: the UNMODIFIED reference classes (EMA, Diffusion.noise_images
from /root/reference/src/diff_modules.py, imported with the matplotlib shim of make_golden.py) run on CPU on seeded
inputs.  Build container only:
    python tests/golden/make_golden_train.py      -> tests/golden/golden_train.npz
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
sys.path.insert(0, HERE)

from make_golden import import_reference  # noqa: E402

EMA_BETA = 0.995
TRAIN_SEED = 77


def train_inputs():
    """(old, new) weights for the EMA average; (x, t) for noise_images at T = 1000."""
    g = torch.Generator(device="cpu")
    g.manual_seed(TRAIN_SEED)
    old = torch.randn(4099, generator=g) * 3
    new = torch.randn(4099, generator=g) * 3
    x = torch.randn(5, 4, 16, 16, generator=g)
    t = torch.tensor([1, 999, 500, 37, 998])
    return old, new, x, t


def main():
    dm, _ = import_reference()
    old, new, x, t = train_inputs()
    out = {}
    ema = dm.EMA(EMA_BETA)
    out["ema_avg"] = ema.update_average(old, new).numpy()
    # step_ema on a tiny module pair: copies while step < step_start_ema, averages afterwards (:42-48)
    torch.manual_seed(TRAIN_SEED)
    model, ema_model = torch.nn.Linear(7, 5), torch.nn.Linear(7, 5)
    w0 = {k: v.clone() for k, v in model.state_dict().items()}
    ema.step_ema(ema_model, model, step_start_ema=1)   # step 0: copy
    assert all(torch.equal(ema_model.state_dict()[k], w0[k]) for k in w0) and ema.step == 1
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.25)
    ema.step_ema(ema_model, model, step_start_ema=1)   # step 1: average
    out["ema_linear_w_start"] = w0["weight"].numpy()
    out["ema_linear_w_after"] = ema_model.state_dict()["weight"].numpy()
    # noise_images (:404-409): the schedule attributes are all the method touches
    beta = torch.linspace(1e-4, 0.02, 1000)
    ns = types.SimpleNamespace(alpha_hat=torch.cumprod(1.0 - beta, dim=0))
    torch.manual_seed(TRAIN_SEED)
    x_t, eps = dm.Diffusion.noise_images(ns, x, t)
    out["noise_x_t"] = x_t.numpy()
    out["noise_eps"] = eps.numpy()
    out["mse"] = np.asarray(torch.nn.MSELoss()(eps, x_t).item(), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "golden_train.npz"), **out)
    print(json.dumps({k: list(v.shape) for k, v in out.items()}))


if __name__ == "__main__":
    main()
