"""Batch-sharded generation over the GPUs of one box.

Every sample's trajectory depends only on its own x_T, label and noise stream, so the sampling loop needs no
data-path collective: the label list is split into contiguous per-rank shards, each rank (one process per
B200) runs its own CUDA-graph loop, and ONE all_gather (NCCL over NVLink) assembles the uint8 output.
The Philox noise stream is keyed by the GLOBAL sample index, so the result does not depend on world size.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous [lo, hi) of rank's shard; the first n % world_size ranks get one extra sample."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_shards(local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """all_gather ragged per-rank shards [n_local, ...] into [n, ...] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    per = max(shard_bounds(n, ws, r)[1] - shard_bounds(n, ws, r)[0] for r in range(ws))
    pad = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]].copy_(local)
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    parts = []
    for r in range(ws):
        lo, hi = shard_bounds(n, ws, r)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts, dim=0)


def sample_sharded(diffusion, labels, cfg_scale=3, *, use_ema=False, seed=0, gather=True, group=None, sample_fn=None,
                   **kw):
    """Diffusion.sample over all ranks: rank r samples labels[lo_r:hi_r] with sample_base = lo_r.

    `sample_fn(labels_shard, sample_base)` replaces the device sampler in host-logic tests (gloo, CPU)."""
    labels = torch.as_tensor(labels).reshape(-1)
    n = len(labels)
    ws = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if ws > 1 else 0
    lo, hi = shard_bounds(n, ws, rank)
    if sample_fn is not None:
        local = sample_fn(labels[lo:hi], lo)
    else:
        local = diffusion.sample(use_ema, labels[lo:hi], cfg_scale, seed=seed, sample_base=lo, **kw)
    return gather_shards(local, n, group) if gather else local
