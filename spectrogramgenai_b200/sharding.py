"""Batch-sharded generation over the GPUs of one box.

Every sample's trajectory depends only on its own x_T, label and noise stream, so the sampling loop needs no
data-path collective: the label list is split into contiguous per-rank shards, each rank (one process per
B200) runs its own CUDA-graph loop, and ONE gather (NCCL over NVLink; to one rank, or all_gather when every rank wants
the result) assembles the uint8 output.
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


def gather_shards(local: torch.Tensor, n: int, group=None, dst=None):
    """Assemble ragged per-rank shards [n_local, ...] into [n, ...].

    dst=None: on every rank (one all_gather).  dst=r: on rank r only (one gather; the other ranks return None and neither
    receive nor copy anything) -- what a generation job wants: the images are consumed in one place, or nowhere at all
    when every rank writes its own files.  The receive buffers are views of the result tensor, so there is no second copy
    when the shards are equal-sized (n % world_size == 0); ragged shards are padded to the largest one and compacted."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n, ws, r)[1] - shard_bounds(n, ws, r)[0] for r in range(ws)]
    per = max(sizes)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} samples, its shard of {n} is {sizes[rank]}")
    send = local.contiguous()
    if send.shape[0] != per:
        send = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]].copy_(local)
    receiver = dst is None or rank == dst
    out = torch.empty((ws * per, *local.shape[1:]), dtype=local.dtype, device=local.device) if receiver else None
    if dst is None:
        dist.all_gather_into_tensor(out, send, group=group)
    else:
        bufs = list(out.view(ws, per, *local.shape[1:]).unbind(0)) if receiver else None
        dist.gather(send, bufs, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    if not receiver:
        return None
    if per * ws == n:
        return out
    return torch.cat([out[r * per: r * per + sizes[r]] for r in range(ws)], dim=0)


def sample_sharded(diffusion, labels, cfg_scale=3, *, use_ema=False, seed=0, gather=True, dst=None, group=None,
                   sample_fn=None, **kw):
    """Diffusion.sample over all ranks: rank r samples labels[lo_r:hi_r] with sample_base = lo_r.

    gather=False returns the local shard; otherwise the shards are assembled on every rank (dst=None) or on rank `dst`
    only (the others return None).  `sample_fn(labels_shard, sample_base)` replaces the device sampler in host-logic
    tests (gloo, CPU)."""
    labels = torch.as_tensor(labels).reshape(-1)
    n = len(labels)
    ws = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if ws > 1 else 0
    lo, hi = shard_bounds(n, ws, rank)
    if sample_fn is not None:
        local = sample_fn(labels[lo:hi], lo)
    else:
        local = diffusion.sample(use_ema, labels[lo:hi], cfg_scale, seed=seed, sample_base=lo, **kw)
    return gather_shards(local, n, group, dst) if gather else local
