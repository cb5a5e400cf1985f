"""Batch sharding for replica parallelism.

The hot path has no cross-sample operation (GroupNorm and attention are per sample), so N GPUs run
N independent replicas: weights are broadcast once from rank 0 (NCCL over NVLink on the GPU box,
gloo in the CPU tests), every batch-major input is sliced per rank, and no collective runs per step.

Classifier-free-guidance tensors are laid out [uncond block | cond block] (reference
stable_diffusion_xl/model.py:109-111): a rank must receive the SAME latents' rows from both blocks.
"""

from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of ``total`` items owned by ``rank`` (remainder spread over the first ranks)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: Tensor, rank: int, world: int) -> Tensor:
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def shard_cfg_batch(x: Tensor, rank: int, world: int) -> Tensor:
    """Slice a [2N, ...] classifier-free-guidance tensor: rows of this rank's latents from the
    unconditional half followed by the same latents' rows from the conditional half."""
    assert x.shape[0] % 2 == 0, "CFG tensors hold an unconditional and a conditional half"
    uncond, cond = x.chunk(2)
    return torch.cat((shard_batch(uncond, rank, world), shard_batch(cond, rank, world)))


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> int:
    """Make every rank's parameters and buffers identical to ``src``'s; returns bytes sent.
    The tensors are overwritten in place under ``no_grad`` (their version counters move, so stamped
    caches notice) and every packed-weight cache of the backend is dropped: a forward that ran before
    the broadcast must not leave packs of the old values behind."""
    sent = 0
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            received = t.detach().clone()
            dist.broadcast(received, src=src)
            t.copy_(received)
            sent += t.numel() * t.element_size()
    from refiners_b200 import backend

    backend.clear_caches()
    return sent


def gather_batch(x: Tensor, world: int, total: int | None = None) -> Tensor:
    """All-gather per-rank results along the batch (end of a run, off the hot path).  ``shard_range``
    spreads a remainder over the first ranks, so shards may differ by one row: every rank pads to the
    largest shard and the padding is cut after the gather.  ``total`` is the unsharded batch size
    (default: ``world`` equal shards)."""
    rank = dist.get_rank()
    total = x.shape[0] * world if total is None else total
    sizes = [hi - lo for lo, hi in (shard_range(total, r, world) for r in range(world))]
    assert x.shape[0] == sizes[rank], f"rank {rank} holds {x.shape[0]} rows, its shard of {total} has {sizes[rank]}"
    widest = max(sizes)
    padded = x.contiguous()
    if padded.shape[0] < widest:
        padded = torch.cat((padded, padded.new_zeros((widest - padded.shape[0], *padded.shape[1:]))))
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([part[:n] for part, n in zip(parts, sizes)])
