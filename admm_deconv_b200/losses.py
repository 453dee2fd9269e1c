"""Host-side mirror of the reference's two training losses over the C ABI of ``include/admmtv_loss.h``
(SURVEY.md section 8 row f-2):

    gmsd(x, y, t=0.0026, α=0.0, reduction=mean) / gmsd_loss          src/metrics/gmsd.jl:13-30   (train.jl:191)
    ssim(x, y, kernel; peakval=1, crop=true, dims=:) / ssim_loss / ssim_loss_fast   src/metrics/ssim.jl:84-164   (train_v2.jl:89)

Same names, argument order and meaning as the reference.  Tensors follow ``ops.py``: a contiguous CUDA fp32
``(B,C,N,M)`` tensor has the memory layout of the Julia ``(M,N,C,B)`` array.  The result is a 0-dim CUDA
tensor.  Gradients flow to BOTH arguments through the hand-written backward kernels: both losses are
symmetric in (x, y), so the pullback w.r.t. the second argument is the same kernels with the images swapped
(it runs only when that argument requires a gradient; the training scripts differentiate the prediction only).
``kernel`` may be any window the reference accepts (ssim.jl:84): None (the 11-tap Gaussian), 1-D taps of a
separable window, a 2-D ``(L2, L1)`` window (separable or not), or a 4-D ``(C, 1, L2, L1)`` per-channel stack;
``crop=False`` pads both images symmetrically first (ssim.jl:104-110).  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
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
        dev = x.device.index or 0
        lb = lossbar.to(torch.float32).contiguous()
        xbar = ybar = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                xbar = torch.empty_like(x)
                lib.gmsd_backward(M, N, C, B, dev, x.data_ptr(), y.data_ptr(), t, alpha, lb.data_ptr(), ws.data_ptr(),
                                  xbar.data_ptr(), _stream())
            if ctx.needs_input_grad[1]:
                # gmsd(x, y) = gmsd(y, x) (gmsd.jl:5-10 is symmetric in the two maps): same kernels, images swapped
                ws2 = _ws(lib.gmsd_workspace_bytes(M, N, C, B), x.device)
                tmp = torch.empty((), dtype=torch.float32, device=x.device)
                ybar = torch.empty_like(y)
                lib.gmsd_forward(M, N, C, B, dev, y.data_ptr(), x.data_ptr(), t, alpha, tmp.data_ptr(), ws2.data_ptr(), _stream())
                lib.gmsd_backward(M, N, C, B, dev, y.data_ptr(), x.data_ptr(), t, alpha, lb.data_ptr(), ws2.data_ptr(),
                                  ybar.data_ptr(), _stream())
        return xbar, ybar, None, None


def gmsd(x: torch.Tensor, y: torch.Tensor, t: float = 0.0026, alpha: float = 0.0, reduction=None) -> torch.Tensor:
    """gmsd.jl:13-27.  ``reduction`` (gmsd.jl:13, default ``Flux.mean``) is applied to the B per-image scores; the kernel
    returns their mean, ``torch.sum`` is that times B (the two reductions ``julia/ADMMTVLosses.jl`` takes as well)."""
    if reduction is None or reduction is torch.mean:
        return _Gmsd.apply(x, y, t, alpha)
    if reduction is torch.sum:
        return _Gmsd.apply(x, y, t, alpha) * float(x.shape[0])
    raise NotImplementedError("gmsd: reduction must be torch.mean (default) or torch.sum")


gmsd_loss = gmsd   # gmsd.jl:30


class _Window:
    """A window of ssim.jl:84 in the form the C ABI takes: ``taps`` (equal separable taps, unrolled kernels) or the
    rank-R factors ``u`` (R x L1, dim 1 = M) and ``v`` (R x L2, dim 2 = N) of an arbitrary 2-D window."""

    def __init__(self, taps=None, u=None, v=None, L1=11, L2=11):
        self.taps, self.u, self.v, self.L1, self.L2 = taps, u, v, L1, L2

    @property
    def general(self) -> bool:
        return self.u is not None


def _window_of(kernel2d: np.ndarray) -> _Window:
    """kernel2d[b, a]: tap at offset a along dim 1 (M, the last tensor axis) and b along dim 2 (N)."""
    W = np.asarray(kernel2d, dtype=np.float64).T          # W[a, b], the Julia (L1, L2) array
    L1, L2 = W.shape
    if not (1 <= L1 <= 11 and 1 <= L2 <= 11):
        raise ValueError("SSIM window must be at most 11 x 11")
    U, S, Vt = np.linalg.svd(W)
    keep = [r for r in range(len(S)) if S[r] > 1e-7 * S[0]] or [0]
    if len(keep) == 1 and L1 == L2:
        # rank one: W = s u v'.  Equal taps in both dimensions take the unrolled kernels.
        su = U[:, 0] * np.sqrt(S[0]); sv = Vt[0] * np.sqrt(S[0])
        if su.sum() < 0:
            su, sv = -su, -sv
        if np.allclose(su, sv, rtol=0, atol=1e-7 * np.abs(su).max()):
            return _Window(taps=[float(t) for t in su], L1=L1, L2=L2)
    u = [[float(U[a, r] * S[r]) for a in range(L1)] for r in keep]
    v = [[float(Vt[r, b]) for b in range(L2)] for r in keep]
    return _Window(u=u, v=v, L1=L1, L2=L2)


def _windows_of(kernel, C: int):
    """kernel argument of ssim -> list of (channel slice, _Window)."""
    if kernel is None:
        return [(slice(0, C), _Window(taps=None))]
    k = kernel.detach().cpu().numpy() if isinstance(kernel, torch.Tensor) else np.asarray(kernel, dtype=np.float64)
    if k.ndim == 1:
        if not 1 <= k.shape[0] <= 11:
            raise ValueError("separable SSIM window must have 1..11 taps")
        return [(slice(0, C), _Window(taps=[float(t) for t in k], L1=k.shape[0], L2=k.shape[0]))]
    if k.ndim == 2:
        return [(slice(0, C), _window_of(k))]
    if k.ndim == 4:   # (Ck, 1, L2, L1), the torch view of the Julia (L1, L2, 1, Ck) array; Ck = 1 is repeated (ssim.jl:96-98)
        if k.shape[1] != 1 or k.shape[0] not in (1, C):
            raise ValueError(f"4-D SSIM window must be ({C} or 1, 1, L2, L1)")
        if k.shape[0] == 1 or all(np.array_equal(k[c], k[0]) for c in range(1, k.shape[0])):
            return [(slice(0, C), _window_of(k[0, 0]))]
        return [(slice(c, c + 1), _window_of(k[c, 0])) for c in range(C)]
    raise ValueError("SSIM window must be None, 1-D taps, a 2-D window or a 4-D (C,1,L2,L1) stack")


class _Ssim(torch.autograd.Function):
    @staticmethod
    def _fwd(lib, x, y, win, peakval, as_loss, need_grad):
        B, C, N, M = x.shape
        dev = x.device.index or 0
        out = torch.empty((), dtype=torch.float32, device=x.device)
        if win.general:
            ws = _ws(lib.ssim_window_workspace_bytes(M, N, C, B, win.L1, win.L2, need_grad), x.device)
            lib.ssim_window_forward(M, N, C, B, dev, x.data_ptr(), y.data_ptr(), win.u, win.v, float(peakval), as_loss,
                                    out.data_ptr(), ws.data_ptr(), need_grad, _stream())
        else:
            ws = _ws(lib.ssim_workspace_bytes(M, N, C, B, win.taps, need_grad), x.device)
            lib.ssim_forward(M, N, C, B, dev, x.data_ptr(), y.data_ptr(), win.taps, float(peakval), as_loss,
                             out.data_ptr(), ws.data_ptr(), need_grad, _stream())
        return out, ws

    @staticmethod
    def _bwd(lib, x, y, win, as_loss, ob, ws):
        B, C, N, M = x.shape
        dev = x.device.index or 0
        xbar = torch.empty_like(x)
        if win.general:
            lib.ssim_window_backward(M, N, C, B, dev, x.data_ptr(), y.data_ptr(), win.u, win.v, as_loss, ob.data_ptr(),
                                     ws.data_ptr(), xbar.data_ptr(), _stream())
        else:
            lib.ssim_backward(M, N, C, B, dev, x.data_ptr(), y.data_ptr(), win.taps, as_loss, ob.data_ptr(), ws.data_ptr(),
                              xbar.data_ptr(), _stream())
        return xbar

    @staticmethod
    def forward(ctx, x, y, win, peakval, as_loss):
        lib = _lib.load()
        _check("x", x); _check("y", y); _check_sizes(x, y)
        x, y = x.contiguous(), y.contiguous()
        need_grad = bool(ctx.needs_input_grad[0])
        with torch.cuda.device(x.device):
            out, ws = _Ssim._fwd(lib, x, y, win, peakval, as_loss, need_grad)
        ctx.save_for_backward(x, y, ws)
        ctx.cfg = (win, float(peakval), as_loss)
        return out

    @staticmethod
    def backward(ctx, outbar):
        lib = _lib.load()
        x, y, ws = ctx.saved_tensors
        win, peakval, as_loss = ctx.cfg
        ob = outbar.to(torch.float32).contiguous()
        xbar = ybar = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                xbar = _Ssim._bwd(lib, x, y, win, as_loss, ob, ws)
            if ctx.needs_input_grad[1]:
                # ssim(x, y) = ssim(y, x) (ssim.jl:112-121 is symmetric): the derivative maps of the swapped call
                _, ws2 = _Ssim._fwd(lib, y, x, win, peakval, as_loss, True)
                ybar = _Ssim._bwd(lib, y, x, win, as_loss, ob, ws2)
        return xbar, ybar, None, None, None


class _PadSymmetric(torch.autograd.Function):
    """NNlib ``pad_symmetric(x, (lo1, hi1, lo2, hi2))`` (ssim.jl:108-109) and its pullback."""

    @staticmethod
    def forward(ctx, x, pads):
        lib = _lib.load()
        _check("x", x)
        x = x.contiguous()
        B, C, N, M = x.shape
        lo1, hi1, lo2, hi2 = pads
        out = torch.empty((B, C, N + lo2 + hi2, M + lo1 + hi1), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            lib.pad_symmetric(M, N, C * B, pads, x.device.index or 0, x.data_ptr(), out.data_ptr(), _stream())
        ctx.cfg = (pads, (B, C, N, M))
        return out

    @staticmethod
    def backward(ctx, outbar):
        lib = _lib.load()
        pads, (B, C, N, M) = ctx.cfg
        ob = outbar.to(torch.float32).contiguous()
        xbar = torch.empty((B, C, N, M), dtype=torch.float32, device=ob.device)
        with torch.cuda.device(ob.device):
            lib.pad_symmetric_adjoint(M, N, C * B, pads, ob.device.index or 0, ob.data_ptr(), xbar.data_ptr(), _stream())
        return xbar, None


def pad_symmetric(x: torch.Tensor, pads) -> torch.Tensor:
    """``pads = (lo1, hi1, lo2, hi2)``: dim 1 is M (the last tensor axis), dim 2 is N."""
    return _PadSymmetric.apply(x, tuple(int(p) for p in pads))


def _ssim_impl(x, y, kernel, peakval, crop, as_loss):
    _check("x", x); _check("y", y); _check_sizes(x, y)
    C = x.shape[1]
    parts = _windows_of(kernel, C)
    vals = []
    for sl, win in parts:
        xs, ys = (x, y) if len(parts) == 1 else (x[:, sl].contiguous(), y[:, sl].contiguous())
        if not crop:
            # ssim.jl:104-110: same-size convolution, padding = (cld(L-1, 2), fld(L-1, 2)) per dimension
            pads = (-(-(win.L1 - 1) // 2), (win.L1 - 1) // 2, -(-(win.L2 - 1) // 2), (win.L2 - 1) // 2)
            xs, ys = pad_symmetric(xs, pads), pad_symmetric(ys, pads)
        vals.append(_Ssim.apply(xs, ys, win, peakval, as_loss))
    # per-channel windows: equal-sized maps, so the mean over (1,2,3) then the batch is the mean of the channel means
    return vals[0] if len(vals) == 1 else torch.stack(vals).mean()


def ssim(x: torch.Tensor, y: torch.Tensor, kernel=None, peakval: float = 1.0, crop: bool = True, dims=None) -> torch.Tensor:
    """ssim.jl:84-124.  ``dims`` is accepted and ignored, as in the reference (its body never reads it)."""
    return _ssim_impl(x, y, kernel, peakval, crop, False)


def ssim_loss(x: torch.Tensor, y: torch.Tensor, kernel=None, peakval: float = 1.0, crop: bool = True, dims=None) -> torch.Tensor:
    """ssim.jl:148  ``1 - ssim(x, y)``"""
    return _ssim_impl(x, y, kernel, peakval, crop, True)


def ssim_loss_fast(x: torch.Tensor, y: torch.Tensor, kernel_length: int = 5, peakval: float = 1.0, crop: bool = True,
                   dims=None) -> torch.Tensor:
    """ssim.jl:160-164: normalised box window of side ``kernel_length``."""
    return ssim_loss(x, y, [1.0 / kernel_length] * kernel_length, peakval, crop)
