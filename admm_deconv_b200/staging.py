"""Batch assembly in front of the ADMM-TV path, over ``include/admmtv_batch.h``.

Mirrors the tail of the reference's input pipeline:

    ImageDataFeeder / get_x_y_images / getindex     src/processing/datafeeder.jl:5-68
    img2tensor (N0f8 -> Float32, channels last)     src/utilities/base_funcs.jl:29-35
    `|> gpu` of the batch                           src/train.jl:50, ToGPU() src/train_v2.jl:60

The reference converts every crop to Float32 on the host, concatenates on dim 4 and uploads 4 bytes per sample.
Here the host only gathers the raw 8-bit crops into a pinned buffer; one kernel on a copy stream converts,
de-interleaves and writes the ``(M,N,C,B)`` batch, double-buffered so that the next batch's upload overlaps the
current step.  Image decoding (Images.load) is out of scope: images are given as decoded ``uint8 (H,W,C)`` arrays.
Multi-GPU: each rank assembles only its contiguous block of the batch (``dist.shard_range``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .dist import shard_range


class ImageDataFeeder:
    """Paired-image dataset with random aligned crops (datafeeder.jl:5-46) and device-side batch assembly."""

    def __init__(self, x_data: Sequence[np.ndarray], y_data: Sequence[np.ndarray], x_shape: Tuple[int, int],
                 y_shape: Tuple[int, int], device="cuda:0", seed: Optional[int] = None, depth: int = 2,
                 rank: int = 0, world: int = 1, resident: bool = False):
        if len(x_data) != len(y_data):
            raise ValueError("x_data and y_data must pair up")
        self.x_data, self.y_data = list(x_data), list(y_data)
        self.x_shape, self.y_shape = tuple(x_shape), tuple(y_shape)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ImageDataFeeder assembles batches on the GPU: there is no CPU fallback")
        self.rng = np.random.default_rng(seed)
        self.depth, self.rank, self.world = int(depth), rank, world
        self._slot = 0
        self._bufs = {}
        self._stream = torch.cuda.Stream(device=self.device)
        self._free = [None] * self.depth   # event: the slot's pinned/device byte buffers were consumed
        # resident=True: the decoded 8-bit dataset is uploaded ONCE and stays in HBM; a batch is then gathered on the device
        # from per-image offsets (admmtv_batch_gather_n0f8) and the host sends 8 bytes per image per step
        self.resident = bool(resident)
        if self.resident:
            self._res = {}
            for which, data in (("x", self.x_data), ("y", self.y_data)):
                widths = {im.shape[1] for im in data}
                chans = {1 if im.ndim == 2 else im.shape[2] for im in data}
                if len(widths) != 1 or len(chans) != 1:
                    raise ValueError("resident=True needs images of one width and channel count (shared strides)")
                starts, total = [], 0
                for im in data:
                    if im.dtype != np.uint8:
                        raise TypeError("images must be uint8 (N0f8)")
                    starts.append(total)
                    total += im.size
                flat = torch.empty(total, dtype=torch.uint8)
                for st, im in zip(starts, data):
                    flat[st:st + im.size] = torch.from_numpy(np.ascontiguousarray(im).reshape(-1))
                self._res[which] = (flat.to(self.device), starts, widths.pop(), chans.pop())

    def __len__(self):                      # datafeeder.jl:49-51
        return len(self.y_data)

    def crop_origin(self, idx: int) -> Tuple[int, int]:
        """datafeeder.jl:43-44 (0-based): one origin shared by the x and y crop."""
        H, W = self.y_data[idx].shape[:2]
        if self.y_shape[0] > H or self.y_shape[1] > W:
            raise ValueError(f"target shape {self.y_shape} exceeds image size {(H, W)}")   # the reference warns and returns
        return int(self.rng.integers(0, H - self.y_shape[0] + 1)), int(self.rng.integers(0, W - self.y_shape[1] + 1))

    def _buffers(self, slot: int, which: str, B: int, shape, C: int):
        key = (slot, which, B, shape, C)
        if key not in self._bufs:
            n = B * shape[0] * shape[1] * C
            self._bufs[key] = (torch.empty(n, dtype=torch.uint8).pin_memory(),
                               torch.empty(n, dtype=torch.uint8, device=self.device))
        return self._bufs[key]

    def getindex(self, idxs: Sequence[int], origins: Optional[List[Tuple[int, int]]] = None):
        """datafeeder.jl:54-68: returns (batch_x, batch_y) as CUDA fp32 ``(B,C,N,M)`` tensors (memory layout of the
        Julia ``(M,N,C,B)`` arrays) holding this rank's block of the batch.  Enqueued on a copy stream; the current
        stream is made to wait for it, so the tensors are safe to use right away."""
        lib = _lib.load()
        idxs = [idxs] if isinstance(idxs, (int, np.integer)) else list(idxs)
        lo, hi = shard_range(len(idxs), self.rank, self.world)
        if origins is None:
            origins = [self.crop_origin(i) for i in idxs]      # every rank draws all origins: same RNG stream everywhere
        idxs, origins = idxs[lo:hi], origins[lo:hi]
        B = len(idxs)
        slot = self._slot
        self._slot = (self._slot + 1) % self.depth
        if self._free[slot] is not None:
            self._free[slot].synchronize()                     # the host may overwrite the pinned buffer again
        out = []
        cur = torch.cuda.current_stream(self.device)
        if self.resident:
            for which, shape in (("x", self.x_shape), ("y", self.y_shape)):
                base, starts, W, C = self._res[which]
                M, N = shape
                offs = torch.tensor([starts[i] + (h0 * W + w0) * C for i, (h0, w0) in zip(idxs, origins)], dtype=torch.int64)
                dst = torch.empty(B, C, N, M, dtype=torch.float32, device=self.device)
                offs_d = offs.to(self.device, non_blocking=True)
                lib.batch_gather_n0f8(M, N, C, B, self.device.index or 0, base.data_ptr(), offs_d.data_ptr(), 1, C * W, C,
                                      dst.data_ptr(), cur.cuda_stream)
                out.append(dst)
            return out[0], out[1]
        for which, data, shape in (("x", self.x_data, self.x_shape), ("y", self.y_data, self.y_shape)):
            C = 1 if data[idxs[0]].ndim == 2 else data[idxs[0]].shape[2]
            M, N = shape
            pin, devb = self._buffers(slot, which, B, shape, C)
            view = pin.numpy().reshape(B, M, N, C)
            for k, (i, (h0, w0)) in enumerate(zip(idxs, origins)):
                img = data[i]
                if img.dtype != np.uint8:
                    raise TypeError("images must be uint8 (N0f8)")
                view[k] = img.reshape(img.shape[0], img.shape[1], C)[h0:h0 + M, w0:w0 + N]
            with torch.cuda.stream(self._stream):
                # allocated ON the copy stream: the caching allocator then never hands out a block that kernels still
                # queued on the compute stream are reading (a block freed there is only reusable there until it drains)
                dst = torch.empty(B, C, N, M, dtype=torch.float32, device=self.device)
                devb.copy_(pin, non_blocking=True)
                # row-major (H,W,C) crops: strides (c, i, j, b) = (1, C*N, C, M*N*C)
                lib.batch_from_n0f8(M, N, C, B, self.device.index or 0, devb.data_ptr(), 1, C * N, C, M * N * C,
                                    dst.data_ptr(), self._stream.cuda_stream)
            dst.record_stream(cur)             # consumed on the compute stream: its later free must wait for that stream
            out.append(dst)
        ev = torch.cuda.Event()
        ev.record(self._stream)
        self._free[slot] = ev
        cur.wait_event(ev)
        return out[0], out[1]

    __getitem__ = getindex
    getobs = getindex                         # MLUtils.getobs, datafeeder.jl:76-78

    def numobs(self):                         # MLUtils.numobs, datafeeder.jl:71-73
        return len(self)

    def h2d_bytes(self, batch: int) -> int:
        if self.resident:
            return 2 * 8 * batch      # one int64 offset per image and array
        cx = 1 if self.x_data[0].ndim == 2 else self.x_data[0].shape[2]
        cy = 1 if self.y_data[0].ndim == 2 else self.y_data[0].shape[2]
        return batch * (self.x_shape[0] * self.x_shape[1] * cx + self.y_shape[0] * self.y_shape[1] * cy)
