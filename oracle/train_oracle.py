"""CPU restatement of the forward-only pieces of the reference's training loop -- TEST INFRASTRUCTURE ONLY
(the backward pass is not built).

Restates
  * EMA.update_average / update_model_average / step_ema   /root/reference/src/diff_modules.py:30-46
  * Diffusion.sample_timesteps / noise_images                :401-409
  * the validation objective of one_epoch(train=False)       :474-478  (noise -> model -> nn.MSELoss)
with the same torch CPU primitives in the same evaluation order (un-fused multiplies / adds), pinned bit-for-bit to the
reference classes by tests/golden/make_golden_train.py -> golden_train.npz.
"""
from __future__ import annotations

import torch


def ema_update_average(old, new, beta):
    """EMA.update_average (:37-40): `old * beta + (1 - beta) * new` -- the Python floats beta and (1 - beta) are
    applied as fp32 scalars; two rounded products, one rounded sum."""
    if old is None:
        return new
    return old * beta + (1 - beta) * new


def ema_step(ema_sd, sd, step, beta, step_start_ema=2000):
    """EMA.step_ema (:42-48) on state dicts: copy while step < step_start_ema, average afterwards.  Returns the new EMA
    state dict (the caller increments step)."""
    if step < step_start_ema:
        return {k: v.clone() for k, v in sd.items()}
    return {k: ema_update_average(ema_sd[k], sd[k], beta) for k in sd}


def noise_images(x, t, alpha_hat, eps):
    """Diffusion.noise_images (:404-409) with the Gaussian draw injected: (x_t, eps)."""
    sqrt_alpha_hat = torch.sqrt(alpha_hat[t])[:, None, None, None]
    sqrt_one_minus_alpha_hat = torch.sqrt(1 - alpha_hat[t])[:, None, None, None]
    return sqrt_alpha_hat * x + sqrt_one_minus_alpha_hat * eps, eps


def mse(noise, pred):
    """nn.MSELoss() (:478): mean over all elements."""
    return torch.nn.functional.mse_loss(noise, pred)
