"""Host-side mirror of the reference's two training losses over the C ABI of ``include/admmtv_loss.h``
(SURVEY.md section 8 row f-2):

    gmsd(x, y, t=0.0026, α=0.0) / gmsd_loss          src/metrics/gmsd.jl:13-30   (train.jl:191)
    ssim(x, y, kernel; peakval=1) / ssim_loss / ssim_loss_fast   src/metrics/ssim.jl:84-164   (train_v2.jl:89)

Same names, argument order and meaning as the reference.  Tensors follow ``ops.py``: a contiguous CUDA fp32
``(B,C,N,M)`` tensor has the memory layout of the Julia ``(M,N,C,B)`` array.  The result is a 0-dim CUDA
tensor; gradients flow to the FIRST argument (the prediction) through the hand-written backward kernels --
the target is treated as a constant, which is how both training scripts use these losses.
There is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib


def _check(name: str, t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the loss kernels have no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    if t.dim() != 4:
        raise ValueError(f"{name} must be 4-D (B,C,N,M)")


def _check_sizes(x: torch.Tensor, y: torch.Tensor):
    """ssim.jl:56-62 _check_sizes"""
    if x.shape != y.shape:
        raise ValueError(f"loss function expects size(ŷ) = {tuple(y.shape)} but is size {tuple(x.shape)}")


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Gmsd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, t, alpha):
        lib = _lib.load()
        _check("x", x); _check("y", y); _check_sizes(x, y)
        x, y = x.contiguous(), y.contiguous()
        B, C, N, M = x.shape
        dev = x.device.index or 0
        ws = _ws(lib.gmsd_workspace_bytes(M, N, C, B), x.device)
        out = torch.empty((), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            lib.gmsd_forward(M, N, C, B, dev, x.data_ptr(), y.data_ptr(), float(t), float(alpha), out.data_ptr(),
                             ws.data_ptr(), _stream())
        ctx.save_for_backward(x, y, ws)
        ctx.cfg = (float(t), float(alpha))
        return out

    @staticmethod
    def backward(ctx, lossbar):
        lib = _lib.load()
        x, y, ws = ctx.saved_tensors
        B, C, N, M = x.shape
        t, alpha = ctx.cfg
        lb = lossbar.to(torch.float32).contiguous()
        xbar = torch.empty_like(x)
        with torch.cuda.device(x.device):
            lib.gmsd_backward(M, N, C, B, x.device.index or 0, x.data_ptr(), y.data_ptr(), t, alpha, lb.data_ptr(),
                              ws.data_ptr(), xbar.data_ptr(), _stream())
        return xbar, None, None, None


def gmsd(x: torch.Tensor, y: torch.Tensor, t: float = 0.0026, alpha: float = 0.0) -> torch.Tensor:
    """gmsd.jl:13-27 with the default ``reduction = mean``."""
    return _Gmsd.apply(x, y, t, alpha)


gmsd_loss = gmsd   # gmsd.jl:30


class _Ssim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, taps, peakval, as_loss):
        lib = _lib.load()
        _check("x", x); _check("y", y); _check_sizes(x, y)
        x, y = x.contiguous(), y.contiguous()
        B, C, N, M = x.shape
        need_grad = bool(ctx.needs_input_grad[0])
        ws = _ws(lib.ssim_workspace_bytes(M, N, C, B, taps, need_grad), x.device)
        out = torch.empty((), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            lib.ssim_forward(M, N, C, B, x.device.index or 0, x.data_ptr(), y.data_ptr(), taps, float(peakval), as_loss,
                             out.data_ptr(), ws.data_ptr(), need_grad, _stream())
        ctx.save_for_backward(x, y, ws)
        ctx.cfg = (taps, as_loss)
        return out

    @staticmethod
    def backward(ctx, outbar):
        lib = _lib.load()
        x, y, ws = ctx.saved_tensors
        B, C, N, M = x.shape
        taps, as_loss = ctx.cfg
        ob = outbar.to(torch.float32).contiguous()
        xbar = torch.empty_like(x)
        with torch.cuda.device(x.device):
            lib.ssim_backward(M, N, C, B, x.device.index or 0, x.data_ptr(), y.data_ptr(), taps, as_loss, ob.data_ptr(),
                              ws.data_ptr(), xbar.data_ptr(), _stream())
        return xbar, None, None, None, None


def _taps_of(kernel) -> Optional[Sequence[float]]:
    """The reference passes a 2-D window (ssim.jl:84 ``kernel_ref``); only separable windows are supported here and
    they are given by their 1-D taps (None = the 11-tap Gaussian of ssim.jl:6-17)."""
    if kernel is None:
        return None
    taps = [float(v) for v in kernel]
    if not 1 <= len(taps) <= 11:
        raise ValueError("separable SSIM window must have 1..11 taps")
    return taps


def ssim(x: torch.Tensor, y: torch.Tensor, kernel=None, peakval: float = 1.0) -> torch.Tensor:
    """ssim.jl:84-124 with ``crop=true, dims=:`` (the defaults)."""
    return _Ssim.apply(x, y, _taps_of(kernel), peakval, False)


def ssim_loss(x: torch.Tensor, y: torch.Tensor, kernel=None, peakval: float = 1.0) -> torch.Tensor:
    """ssim.jl:148  ``1 - ssim(x, y)``"""
    return _Ssim.apply(x, y, _taps_of(kernel), peakval, True)


def ssim_loss_fast(x: torch.Tensor, y: torch.Tensor, kernel_length: int = 5, peakval: float = 1.0) -> torch.Tensor:
    """ssim.jl:160-164: normalised box window of side ``kernel_length``."""
    return ssim_loss(x, y, [1.0 / kernel_length] * kernel_length, peakval)
