"""`set_seed` of the reference (/root/reference/src/diff_utils.py:15-32): seeds every generator the
sampling path can touch.  The reference draws x_T on the CPU generator (:418) and the per-step noise on
the device generator (:433); this implementation draws both either from injected tensors or from its own
counter-based Philox stream keyed by `seed`, so the value set here is also stored for `Diffusion.sample`."""
from __future__ import annotations

import os
import random

import numpy as np
import torch

_last_seed = 0


def set_seed(s, reproducible=False):
    "Set random seed for `random`, `torch`, and `numpy` (where available)"
    global _last_seed
    _last_seed = int(s)
    try:
        torch.manual_seed(s)
    except NameError:
        pass
    try:
        torch.cuda.manual_seed_all(s)
    except (NameError, RuntimeError):
        pass
    try:
        np.random.seed(s % (2**32 - 1))
    except NameError:
        pass
    random.seed(s)
    if reproducible:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
        os.environ.setdefault("PYTHONHASHSEED", str(s))


def last_seed() -> int:
    return _last_seed
