"""Multi-GPU plumbing for the ADMM-TV layer.

The path shards naturally: every (channel, image) plane is independent (anisotropic), so the batch
dimension is split into contiguous blocks of whole images -- contiguous memory in the (M,N,P,B)
layout -- one process per GPU, with the PSF / λ / ρ replicated.  The forward needs NO collective.
The backward produces per-rank partial parameter gradients [hbar (kh*kw), lambar, rhobar, biasbar];
they are packed into ONE buffer and summed with ONE all-reduce (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The message is < 4 KB: latency-bound, never bandwidth-bound.

Isotropic TV couples the planes of a call through the per-pixel norm; sharded, each rank uses its
own shard's norm ("per-shard batch" semantics == running the reference on each shard).
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the batch for `rank`; sizes differ by at most one image."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(y: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """y is (B,P,N,M); returns this rank's contiguous block of whole images (a view)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(y.shape[0], rank, world)
    return y[lo:hi]


def pack_grads(params: Iterable[torch.nn.Parameter]) -> torch.Tensor:
    parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params]
    return torch.cat(parts) if parts else torch.empty(0)


def unpack_grads(buf: torch.Tensor, params: Iterable[torch.nn.Parameter]) -> None:
    o = 0
    for p in params:
        n = p.numel()
        p.grad = buf[o:o + n].reshape(p.shape).clone()
        o += n


def allreduce_layer_grads(layers, group=None, average: bool = False) -> int:
    """Sum (or average) the trainable parameters' gradients of one or several ADMM layers across
    ranks with a single all-reduce.  Returns the number of floats sent."""
    if not isinstance(layers, (list, tuple)):
        layers = [layers]
    params = [p for l in layers for p in l.parameters() if p.requires_grad and p.numel() > 0]
    buf = pack_grads(params)
    if buf.numel() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if average:
            buf /= dist.get_world_size(group)
    unpack_grads(buf, params)
    return int(buf.numel())
