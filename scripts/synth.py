"""Synthetic checkpoints in the reference's file layout (no network, no shipped weights): shared by the job scripts."""
import os

import torch


def synthetic_vqae(seed=0):
    """Random VQAE decoder + codebook of the reference's shapes (diff_modules.py:266-270, :326-334)."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *shape, fan: (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / fan) ** 0.5  # noqa: E731
    return {"codebook.embedding": torch.rand(512, 4, generator=g) * 2 - 1,
            "decoder.in_proj.weight": u(512, 4, 1, 1, fan=4), "decoder.in_proj.bias": u(512, fan=300),
            "decoder.residual_conv_1.weight": u(512, 512, 1, 1, fan=512), "decoder.residual_conv_1.bias": u(512, fan=300),
            "decoder.residual_conv_2.weight": u(512, 512, 3, 3, fan=4608), "decoder.residual_conv_2.bias": u(512, fan=300),
            "decoder.strided_t_conv_1.weight": u(512, 512, 2, 2, fan=512), "decoder.strided_t_conv_1.bias": u(512, fan=300),
            "decoder.strided_t_conv_2.weight": u(512, 1, 2, 2, fan=512), "decoder.strided_t_conv_2.bias": u(1, fan=300)}


def write_generation_inputs(work, num_classes=27, seed=42):
    """models/<run>/ckpt.pt + optim.pt, models/VQAE/ckpt.pt and data/train/<class dirs> under `work` -- what
    ddpm_conditional_generate.py (and spectrogramgenai_b200.generate) expect to find on disk."""
    from spectrogramgenai_b200.diff_modules import UNet_conditional

    run = os.path.join(work, "models", "DDPM_conditional_VAE")
    os.makedirs(run, exist_ok=True)
    os.makedirs(os.path.join(work, "models", "VQAE"), exist_ok=True)
    for k in range(num_classes):
        os.makedirs(os.path.join(work, "data", "train", f"class{k:02d}"), exist_ok=True)
    torch.manual_seed(seed)
    m = UNet_conditional(4, 4, num_classes=num_classes)
    torch.save(m.state_dict(), os.path.join(run, "ckpt.pt"))
    torch.save({}, os.path.join(run, "optim.pt"))
    torch.save(synthetic_vqae(), os.path.join(work, "models", "VQAE", "ckpt.pt"))
