"""Host-buffer calls (include/admmtv_host.h): what a caller whose arrays live in CPU memory uses --
the reference's ``tvd_fft`` on a CPU ``Array`` (src/ops/ops.jl:183-187) and one step of ``train.jl:49-54``
(batch from the DataLoader |> gpu, withgradient, gradients back).  Everything still runs on the GPU
through libadmmtv.so; this module only owns the session handle, its device arena (a torch uint8 tensor, so
torch's allocator accounts for it) and the compute stream.

Arrays are contiguous CPU ``float32`` tensors in the (B,P,N,M) layout of ``ops.py`` (== Julia (M,N,P,B)
column-major).  Pinned tensors transfer asynchronously at full PCIe speed; pageable ones work but block."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def _hptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if t.is_cuda:
        raise RuntimeError("host session arguments must be CPU tensors (device tensors go through ops.tvd_fft)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("host session arguments must be contiguous float32")
    return t.data_ptr()


class HostSession:
    """Two-slot pipelined host-buffer session: ``enqueue`` step i+1 before ``wait``-ing for step i and the
    host->device copy of the next batch overlaps the kernels of the current one."""

    def __init__(self, M, N, P, B, kh=0, kw=0, iters=10, iso=False, activation="identity", has_bias=False, device=0,
                 flags=0, creg=0.0, groups=0, training=False, group=None):
        self.lib = _lib.load()
        self.desc = _lib.make_desc(M, N, P, B, kh, kw, iters, iso, activation, has_bias, device, flags, creg, groups)
        self.training = bool(training)
        self.device = torch.device("cuda", device)
        nbytes = self.lib.host_session_bytes(self.desc, self.training)
        self._arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._stream = torch.cuda.Stream(device=self.device)
        self._sess = self.lib.host_session_create(self.desc, self.training, self._arena.data_ptr(), self._stream.cuda_stream)
        self.ngrad = self.lib.host_grad_floats(self.desc)
        self._hooks = None
        self._group = group
        import torch.distributed as dist
        if self.training and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self._cb = _lib.ALLREDUCE_FN(self._allreduce)      # keep the ctypes thunk alive
            self._hooks = _lib.Hooks(self._cb, None, 1)

    # data-parallel gradient all-reduce (NCCL over NVLink), stream-ordered on the session's compute stream
    def _allreduce(self, buf, count, stream, user) -> int:
        try:
            import torch.distributed as dist
            off = int(buf) - self._arena.data_ptr()
            view = self._arena[off:off + 4 * int(count)].view(torch.float32)
            with torch.cuda.stream(self._stream):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self._group)
            return 0
        except Exception as e:   # never let an exception cross the C boundary
            import sys
            print(f"HostSession all-reduce failed: {e!r}", file=sys.stderr)
            return -8

    def launches(self) -> int:
        return self.lib.host_launches(self._sess, self.training)

    def forward_enqueue(self, slot, y, lam, rho, h=None, bias=None, out=None):
        out = torch.empty_like(y) if out is None else out
        self.lib.host_forward_enqueue(self._sess, slot, _hptr(y), _hptr(h), _hptr(lam), _hptr(rho), _hptr(bias), _hptr(out))
        return out

    def train_step_enqueue(self, slot, y, target, lam, rho, h=None, bias=None, grads=None, loss=None, ybar=None):
        """grads: CPU float32 [ngrad] = [hbar | lambdabar | rhobar | biasbar]; loss: CPU float32 [1]."""
        grads = torch.empty(self.ngrad, dtype=torch.float32) if grads is None else grads
        loss = torch.empty(1, dtype=torch.float32) if loss is None else loss
        self.lib.host_train_step_enqueue(self._sess, slot, _hptr(y), _hptr(target), _hptr(h), _hptr(lam), _hptr(rho),
                                         _hptr(bias), _hptr(grads), _hptr(loss), _hptr(ybar), self._hooks)
        return grads, loss

    def _n0f8_strides(self, layout):
        M, N, P = self.desc.M, self.desc.N, self.desc.P
        if layout == "BCNM":
            return (M * N, 1, M, M * N * P)          # channel, dim 1 (M), dim 2 (N), image
        if layout == "BNMC":
            return (1, P, P * M, M * N * P)
        raise ValueError("layout must be 'BCNM' or 'BNMC'")

    def forward_enqueue_n0f8(self, slot, y_u8, lam, rho, h=None, bias=None, out=None, layout="BCNM"):
        """Forward fed with 8-bit samples (value / 255); ``out`` is the fp32 (B,P,N,M) result (``admmtv_host_forward_enqueue_n0f8``)."""
        if y_u8.is_cuda or y_u8.dtype != torch.uint8 or not y_u8.is_contiguous():
            raise TypeError("n0f8 host arguments must be contiguous CPU uint8 tensors")
        d = self.desc
        out = torch.empty((d.B, d.P, d.N, d.M), dtype=torch.float32) if out is None else out
        self.lib.host_forward_enqueue_n0f8(self._sess, slot, y_u8.data_ptr(), self._n0f8_strides(layout), _hptr(h), _hptr(lam),
                                           _hptr(rho), _hptr(bias), _hptr(out))
        return out

    def train_step_enqueue_n0f8(self, slot, y_u8, target_u8, lam, rho, h=None, bias=None, grads=None, loss=None, layout="BCNM"):
        """The same step fed with 8-bit samples (value / 255), the dataset's own format: 1 byte per sample crosses PCIe and the
        conversion to the fp32 (M,N,P,B) batch runs on the device (``admmtv_host_train_step_enqueue_n0f8``).
        ``layout``: "BCNM" = planar uint8 (B,P,N,M) like the float tensors; "BNMC" = channel-interleaved images (B,N,M,P) as
        decoded image files are (``staging.ImageDataFeeder``'s convention)."""
        for t in (y_u8, target_u8):
            if t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous():
                raise TypeError("n0f8 host arguments must be contiguous CPU uint8 tensors")
        strides = self._n0f8_strides(layout)
        grads = torch.empty(self.ngrad, dtype=torch.float32) if grads is None else grads
        loss = torch.empty(1, dtype=torch.float32) if loss is None else loss
        self.lib.host_train_step_enqueue_n0f8(self._sess, slot, y_u8.data_ptr(), target_u8.data_ptr(), strides, _hptr(h), _hptr(lam),
                                              _hptr(rho), _hptr(bias), _hptr(grads), _hptr(loss), self._hooks)
        return grads, loss

    def wait(self, slot):
        self.lib.host_wait(self._sess, slot)

    def forward(self, y, lam, rho, h=None, bias=None, out=None):
        out = self.forward_enqueue(0, y, lam, rho, h, bias, out)
        self.wait(0)
        return out

    def train_step(self, y, target, lam, rho, h=None, bias=None, ybar=None):
        g, l = self.train_step_enqueue(0, y, target, lam, rho, h, bias, ybar=ybar)
        self.wait(0)
        return g, l

    def close(self):
        if getattr(self, "_sess", None):
            self.lib.host_session_destroy(self._sess)
            self._sess = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
