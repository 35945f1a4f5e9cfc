"""GPU parity of the forward-only training helpers (sg_noise_images, sg_ema_update, sg_mse) with the reference:
golden outputs of the unmodified reference classes (tests/golden/golden_train.npz) and the CPU oracle
(oracle/train_oracle.py).  Bars: noise_images and the EMA average bit-exact (un-fused fp32 arithmetic in the
reference's order); MSE within 1e-6 relative (different summation order, double partial sums)."""
import os

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import train_oracle as TO
from oracle.weights import make_state_dict
from tests.golden.make_golden_train import EMA_BETA, train_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_train.npz"))


def _diffusion(T=1000, s=16, mode="fp32"):
    from spectrogramgenai_b200.diff_modules import Diffusion

    d = Diffusion(noise_steps=T, img_size=s, num_classes=27, c_in=4, c_out=4, device=DEV, compute_dtype=mode)
    d.model.load_state_dict(make_state_dict(1234, 4, 4, 27))
    return d


def test_noise_images_matches_reference_bit_exactly():
    d = _diffusion()
    _, _, x, t = train_inputs()
    x_t, eps = d.noise_images(x, t, noise=torch.from_numpy(G["noise_eps"]))
    assert np.array_equal(x_t.cpu().numpy(), G["noise_x_t"])
    assert np.array_equal(eps.cpu().numpy(), G["noise_eps"])
    # Philox draw: standard normal, reproducible, keyed by the global sample index (split-invariant)
    big = torch.zeros(64, 4, 16, 16)
    tt = torch.full((64,), 999)
    a, ea = d.noise_images(big, tt, seed=5)
    b, eb = d.noise_images(big[32:], tt[32:], seed=5, sample_base=32)
    assert torch.equal(ea[32:], eb) and torch.equal(a[32:], b)
    assert abs(float(ea.mean())) < 0.02 and abs(float(ea.std()) - 1.0) < 0.02
    c, ec = d.noise_images(big, tt, seed=6)
    assert not torch.equal(ea, ec)
    assert d.sample_timesteps(1000).min() >= 1 and d.sample_timesteps(1000).max() <= 999


def test_ema_average_and_step_match_reference():
    from spectrogramgenai_b200 import ops
    from spectrogramgenai_b200.diff_modules import EMA, UNet_conditional

    old, new, _, _ = train_inputs()
    got = ops.ema_update(old.to(DEV).clone(), new.to(DEV), EMA_BETA)
    assert np.array_equal(got.cpu().numpy(), G["ema_avg"])
    # step_ema on two UNets: copy before step_start_ema, average afterwards -- against the oracle on the state dicts
    sd_a, sd_b = make_state_dict(1, 4, 4, 27), make_state_dict(2, 4, 4, 27)
    model = UNet_conditional(4, 4, num_classes=27, compute_dtype="fp32").to(DEV)
    ema_model = UNet_conditional(4, 4, num_classes=27, compute_dtype="fp32").to(DEV)
    model.load_state_dict(sd_a)
    ema = EMA(EMA_BETA)
    ema.step_ema(ema_model, model, step_start_ema=1)
    assert all(torch.equal(v.cpu(), sd_a[k]) for k, v in ema_model.state_dict().items()) and ema.step == 1
    model.load_state_dict(sd_b)
    ema.step_ema(ema_model, model, step_start_ema=1)
    want = TO.ema_step(sd_a, sd_b, step=1, beta=EMA_BETA, step_start_ema=1)
    for k, v in ema_model.state_dict().items():
        assert torch.equal(v.cpu(), want[k]), k
    assert ema.step == 2
    # the averaged weights are what the next forward uses (packed weights are rebuilt)
    x = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    t = torch.tensor([500, 20])
    eps = ema_model(x.to(DEV), t.to(DEV), torch.tensor([1, 2], device=DEV))
    ref = O.unet_forward(want, x, t, torch.tensor([1, 2]))
    assert O.rel_l2(eps.cpu(), ref) < 1e-4


@pytest.mark.parametrize("n", [1, 1000, 4 * 16 * 16 * 5, 3_000_001])
def test_mse_matches_torch(n):
    from spectrogramgenai_b200 import ops

    g = torch.Generator().manual_seed(n)
    a, b = torch.randn(n, generator=g), torch.randn(n, generator=g) * 2
    got = float(ops.mse(a.to(DEV), b.to(DEV)))
    want = float(((a.double() - b.double()) ** 2).mean())
    assert abs(got - want) <= 1e-6 * want + 1e-12
    assert float(ops.mse(a.to(DEV), b.to(DEV))) == got  # deterministic


def test_eval_loss_is_the_validation_objective():
    """eval_loss = one_epoch(train=False) for one batch (:474-478): noise_images -> model -> MSE, against the oracle
    with the same timesteps and noise (fp32 engine)."""
    d = _diffusion()
    _, _, x, t = train_inputs()
    eps = torch.from_numpy(G["noise_eps"])
    labels = torch.tensor([3, 0, 26, 7, 7])
    got = float(d.eval_loss(x, labels.to(DEV), t=t, noise=eps))
    _, _, alpha_hat = O.noise_schedule(1000, 1e-4, 0.02)
    x_t, _ = TO.noise_images(x, t, alpha_hat, eps)
    pred = O.unet_forward(make_state_dict(1234, 4, 4, 27), x_t, t, labels)
    want = float(TO.mse(eps, pred))
    assert abs(got - want) <= 1e-4 * want
    un = float(d.eval_loss(x, None, t=t, noise=eps))  # the unconditional branch (labels = None, :475-476)
    want_un = float(TO.mse(eps, O.unet_forward(make_state_dict(1234, 4, 4, 27), x_t, t, None)))
    assert abs(un - want_un) <= 1e-4 * want_un
