"""Multi-GPU plumbing for the ADMM-TV layer.

The path shards naturally: every (channel, image) plane is independent (anisotropic), so the batch
dimension is split into contiguous blocks of whole images -- contiguous memory in the (M,N,P,B)
layout -- one process per GPU, with the PSF / λ / ρ replicated.  The forward needs NO collective.
The backward produces per-rank partial parameter gradients [hbar (kh*kw), lambar, rhobar, biasbar];
they are packed into ONE buffer and summed with ONE all-reduce (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The message is < 4 KB: latency-bound, never bandwidth-bound.

Isotropic TV couples the planes of a call through the per-pixel norm; sharded, each rank uses its
own shard's norm by default ("per-shard batch" semantics == running the reference on each shard).
``IsoCoupling`` restores the reference's single-device semantics (one norm per pixel over the WHOLE batch,
ops.jl:6,10) across ranks: the library calls back once per iteration with the per-pixel partial sums and
this module all-reduces them -- an (M,N) float image per iteration, 1 MB at 512^2.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


class IsoCoupling:
    """Builds the ``admmtv_hooks`` (include/admmtv.h) whose ``allreduce_sum`` sums a device buffer over ``group``.

    The callback receives a raw pointer into one of the caller-owned buffers (workspace / checkpoint);
    ``register`` tells the object about those buffers so the pointer can be turned back into a tensor view.
    torch.distributed collectives are stream-ordered with respect to the current stream, which is the stream the
    library launches on, so no host synchronisation is added."""

    def __init__(self, group=None):
        from . import _lib
        self.group = group
        self._bufs = []
        self.calls = 0
        self.floats = 0
        self._cb = _lib.ALLREDUCE_FN(self._allreduce)       # keep the ctypes thunk alive
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.hooks = _lib.Hooks(self._cb, None, 1 if rank == 0 else 0)

    def register(self, *tensors):
        self._bufs = [t for t in tensors if t is not None]

    def _view(self, ptr: int, count: int) -> torch.Tensor:
        for t in self._bufs:
            base = t.data_ptr()
            if base <= ptr and ptr + 4 * count <= base + t.numel() * t.element_size():
                off = ptr - base
                return t.view(torch.uint8).reshape(-1)[off:off + 4 * count].view(torch.float32)
        raise RuntimeError("IsoCoupling: pointer outside the registered buffers")

    def _allreduce(self, buf, count, stream, user) -> int:
        try:
            dist.all_reduce(self._view(int(buf), int(count)), op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
            self.floats += int(count)
            return 0
        except Exception as e:   # never let an exception cross the C boundary
            import sys
            print(f"IsoCoupling all-reduce failed: {e!r}", file=sys.stderr)
            return -8


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
