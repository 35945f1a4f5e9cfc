"""Seeding helper with the reference's name and call signature (`diff_utils.set_seed(s, reproducible=False)`,
/root/reference/src/diff_utils.py:15-32; called by ddpm_conditional_generate.py:94 before sampling).

The reference draws x_T on the CPU generator (diff_modules.py:418) and the per-step noise on the device generator
(:433).  This package draws both from injected tensors or from its own counter-based Philox stream, which is keyed by an
explicit `seed` argument of `Diffusion.sample`; `set_seed` therefore seeds the host generators (python, numpy, torch CPU
and CUDA) for callers that build inputs with them and remembers the value (`last_seed()`) for drivers that want to pass
it on."""
from __future__ import annotations

import random

import numpy as np
import torch

_SEED = {"value": 0}


def set_seed(s, reproducible=False):
    """Seed python's, numpy's and torch's generators with `s`.  `reproducible` is accepted for signature compatibility:
    the kernels of this package are deterministic by construction (fixed reduction orders, no atomics, no cuDNN), so there
    is no autotuner or non-deterministic algorithm to switch off."""
    s = int(s)
    _SEED["value"] = s
    random.seed(s)
    np.random.seed(s % (2**32 - 1))  # numpy accepts 32-bit seeds only; same reduction as the reference
    torch.manual_seed(s)  # seeds the CPU generator and (lazily) every CUDA generator
    return s


def last_seed() -> int:
    """The value of the most recent set_seed() call (0 before any)."""
    return _SEED["value"]
