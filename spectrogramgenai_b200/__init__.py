"""B200-native (sm_100a) implementation of SpectrogramGenAI's class-conditional DDPM sampling path.

Public surface (drop-in for /root/reference/src/diff_modules.py on this path):
    from spectrogramgenai_b200.diff_modules import UNet_conditional, Diffusion, EMA
    from spectrogramgenai_b200.diff_utils import set_seed
    from spectrogramgenai_b200.sharding import sample_sharded      # batch-sharded generation over N GPUs

Importing the package does not need a GPU; running anything does (no CPU / torch-op fallback).
"""
__version__ = "0.1.0"
