"""Host-side mirror of the reference operator ``tvd_fft`` (src/ops/ops.jl:181-188) over the C ABI.

Array convention.  The reference's arrays are Julia ``(M,N,P,B)`` column-major.  A contiguous
torch tensor of shape ``(B,P,N,M)`` has exactly that memory layout, so that is what this module
takes and returns ("Julia index order reversed"); the PSF ``h`` (Julia ``(kh,kw,1,1)``) is a
contiguous ``(kw,kh)`` / ``(1,1,kw,kh)`` tensor.  ``to_julia`` / ``from_julia`` convert to the
``(M,N,P,B)``-indexed view the oracle uses.

PyTorch is used for device memory, streams and autograd plumbing only; all arithmetic happens in
libadmmtv.so (hand-written sm_100a kernels).  There is no fallback: a missing library or a CPU
tensor raises.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def to_julia(t: torch.Tensor) -> torch.Tensor:
    """(B,P,N,M) -> (M,N,P,B)-indexed view (no copy)."""
    return t.permute(*reversed(range(t.dim())))


def from_julia(t: torch.Tensor) -> torch.Tensor:
    """(M,N,P,B)-indexed array -> contiguous (B,P,N,M) tensor."""
    return t.permute(*reversed(range(t.dim()))).contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _check_cuda_f32(name: str, t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the ADMM-TV path has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _kernel_dims(h: Optional[torch.Tensor]):
    if h is None or h.numel() == 0:
        return 0, 0
    if h.dim() == 4:
        return int(h.shape[3]), int(h.shape[2])
    if h.dim() == 2:
        return int(h.shape[1]), int(h.shape[0])
    raise ValueError("h must be (kw,kh) or (1,1,kw,kh)")


def make_desc_for(y: torch.Tensor, h, iters: int, iso: bool, activation="identity", has_bias=False, flags=0,
                  creg=0.0) -> _lib.Desc:
    B, P, N, M = y.shape
    kh, kw = _kernel_dims(h)
    return _lib.make_desc(M, N, P, B, kh, kw, iters, iso, activation, has_bias, y.device.index or 0, flags, creg)


def _alloc(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns 512-byte aligned blocks
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)


class _AdmmFunction(torch.autograd.Function):
    """forward = admmtv_forward, backward = admmtv_backward (the hand-written adjoint that replaces
    Zygote's tape through ops.jl:166-174)."""

    @staticmethod
    def forward(ctx, y, lam, rho, h, bias, iters, iso, activation, creg, flags, coupling=None, need_grad=True):
        lib = _lib.load()
        _check_cuda_f32("y", y)
        for n, t in (("lambda", lam), ("rho", rho)):
            _check_cuda_f32(n, t)
        if h is not None and h.numel() > 0:
            _check_cuda_f32("h", h)
        else:
            h = None
        if bias is not None:
            _check_cuda_f32("bias", bias)
        d = make_desc_for(y, h, iters, iso, activation, bias is not None, flags, creg)
        fwd_b, ck_b, _ = lib.workspace_bytes(d)
        ws = _alloc(fwd_b, y.device)
        ck = _alloc(ck_b, y.device) if need_grad else None
        x = torch.empty_like(y)
        stream = torch.cuda.current_stream(y.device).cuda_stream
        # lam / rho / h are clamped IN PLACE (deconv_admm.jl:216-219 persists the clamp)
        if coupling is not None and iso:
            coupling.register(ws, ck)
            lib.forward_ex(d, _ptr(y), _ptr(h), _ptr(lam), _ptr(rho), _ptr(bias), _ptr(x), _ptr(ws), _ptr(ck), stream,
                           coupling.hooks)
        else:
            coupling = None
            lib.forward(d, _ptr(y), _ptr(h), _ptr(lam), _ptr(rho), _ptr(bias), _ptr(x), _ptr(ws), _ptr(ck), stream)
        ctx.coupling = coupling
        ctx.desc = d
        ctx.ck = ck
        ctx.has_h = h is not None
        ctx.has_bias = bias is not None
        ctx.save_for_backward(y, lam, rho, h if h is not None else torch.empty(0, device=y.device), x)
        return x

    @staticmethod
    def backward(ctx, xbar):
        lib = _lib.load()
        y, lam, rho, h, x = ctx.saved_tensors
        d = ctx.desc
        xbar = xbar.contiguous()
        _, _, bwd_b = lib.workspace_bytes(d)
        ws = _alloc(bwd_b, y.device)
        ybar = torch.empty_like(y)
        hbar = torch.empty_like(h) if ctx.has_h else None
        lbar = torch.empty_like(lam)
        rbar = torch.empty_like(rho)
        bbar = torch.empty(1, dtype=torch.float32, device=y.device) if ctx.has_bias else None
        stream = torch.cuda.current_stream(y.device).cuda_stream
        args = (d, _ptr(xbar), _ptr(x), _ptr(y), _ptr(h) if ctx.has_h else None, _ptr(lam), _ptr(rho), _ptr(ctx.ck),
                _ptr(ybar), _ptr(hbar), _ptr(lbar), _ptr(rbar), _ptr(bbar), _ptr(ws), stream)
        if ctx.coupling is not None:
            ctx.coupling.register(ws, ctx.ck)
            lib.backward_ex(*args, ctx.coupling.hooks)
        else:
            lib.backward(*args)
        return ybar, lbar, rbar, hbar, bbar, None, None, None, None, None, None, None


def _needs_grad(*tensors) -> bool:
    """Evaluated OUTSIDE autograd.Function.forward (grad mode is always off in there): under torch.no_grad() / eval
    nothing is differentiated, so no per-iteration checkpoint is allocated or written (it is (3K-1)*8 bytes per
    pair-pixel: 60 GB at 64 x 512^2 x 3, K = 100) and the faster inference kernels run."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _own_if_nonleaf(t, clamp: bool):
    """The persisted clamp (deconv_admm.jl:216-219) writes into the parameter buffers.  That is the reference's
    behaviour for the layer's own arrays (leaf tensors); a NON-leaf input (e.g. softplus(raw), a torch.cat of several
    layers' scalars) may have been saved by the node that produced it, so the kernel gets a private differentiable
    copy to clamp instead of mutating it behind autograd's back."""
    if t is None or not clamp or t.grad_fn is None:
        return t
    return t.clone()


def admm_layer_call(y, lam, rho, h=None, bias=None, iters=100, iso=False, activation="identity", creg=0.0,
                    nograd_repeat=False, clamp=True, iso_coupling=None):
    """The full layer call (d::Admm)(x), deconv_admm.jl:215-225, differentiable.

    ``iso_coupling`` (a ``dist.IsoCoupling``, EXTENSION): with a batch sharded over ranks, makes the
    isotropic per-pixel norm span every rank's images, i.e. the reference's single-device result on the whole batch."""
    flags = (0 if clamp else _lib.FLAG_NO_CLAMP) | (_lib.FLAG_NOGRAD_REPEAT if nograd_repeat else 0)
    if lam.numel() > 1 or rho.numel() > 1:
        # EXTENSION: one (lambda_k, rho_k) per unrolled iteration (ADMMTV_FLAG_PER_ITER_PARAMS, include/admmtv.h)
        if lam.numel() != int(iters) or rho.numel() != int(iters):
            raise ValueError("per-iteration parameters: lam and rho need exactly `iters` values each")
        flags |= _lib.FLAG_PER_ITER_PARAMS
    lam, rho, h = _own_if_nonleaf(lam, clamp), _own_if_nonleaf(rho, clamp), _own_if_nonleaf(h, clamp)
    return _AdmmFunction.apply(y, lam, rho, h, bias, int(iters), bool(iso), activation, float(creg), flags, iso_coupling,
                               _needs_grad(y, lam, rho, h, bias))


def tvd_fft(y, lam, rho, h=None, isotropic=False, maxit=100, iso_coupling=None):
    """tvd_fft(y, λ, ρ, h, isotropic, maxit) -- ops.jl:181: no clamp, no bias, no activation."""
    return admm_layer_call(y, lam, rho, h, None, maxit, isotropic, "identity", 0.0, False, clamp=False,
                           iso_coupling=iso_coupling)


tvd_fft_gpu = tvd_fft  # ops.jl:99 (tests/admm_deconv_test.jl:76 calls it directly)


def tvd_fft_host(y, lam: float, rho: float, h=None, isotropic=False, maxit=100):
    """Host-buffer call (the reference's tvd_fft on a CPU Array): numpy in, numpy out, the copies
    happen inside the C ABI (admmtv_forward_host).  Still runs on the GPU -- no CPU arithmetic."""
    import numpy as np

    lib = _lib.load()
    y = np.ascontiguousarray(y, dtype=np.float32)  # (B,P,N,M)
    B, P, N, M = y.shape
    kh, kw = (0, 0) if h is None else (h.shape[-1], h.shape[-2])
    d = _lib.make_desc(M, N, P, B, kh, kw, maxit, isotropic, "identity", False, 0, _lib.FLAG_NO_CLAMP, 0.0)
    hb = None if h is None else np.ascontiguousarray(h, dtype=np.float32)
    lb = np.array([lam], dtype=np.float32)
    rb = np.array([rho], dtype=np.float32)
    x = np.empty_like(y)
    lib.forward_host(d, y.ctypes.data, None if hb is None else hb.ctypes.data, lb.ctypes.data, rb.ctypes.data, None,
                     x.ctypes.data)
    return x


class _AdmmGroupedFunction(torch.autograd.Function):
    """Grouped call (desc.groups > 1): forward = admmtv_forward, backward = admmtv_backward with per-group
    parameter cotangents (and ybar summed over groups that share the input)."""

    @staticmethod
    def forward(ctx, y, lam, rho, h, bias, iters, iso, activation, creg, flags, groups, shared_input, channel_concat,
                need_grad=True):
        lib = _lib.load()
        Bin, P, N, M = y.shape
        Bg = Bin if shared_input else Bin // groups
        kh, kw = (0, 0) if h is None else (int(h.shape[-1]), int(h.shape[-2]))
        d = _lib.make_desc(M, N, P, groups * Bg, kh, kw, iters, iso, activation, bias is not None, y.device.index or 0,
                           flags, creg, groups)
        fwd_b, ck_b, _ = lib.workspace_bytes(d)
        ws = _alloc(fwd_b, y.device)
        ck = _alloc(ck_b, y.device) if need_grad else None
        x = torch.empty((Bg, groups * P, N, M) if channel_concat else (groups * Bg, P, N, M), dtype=torch.float32,
                        device=y.device)
        stream = torch.cuda.current_stream(y.device).cuda_stream
        lib.forward(d, _ptr(y), _ptr(h), _ptr(lam), _ptr(rho), _ptr(bias), _ptr(x), _ptr(ws), _ptr(ck), stream)
        ctx.desc, ctx.ck, ctx.has_h, ctx.has_bias, ctx.groups = d, ck, h is not None, bias is not None, groups
        ctx.save_for_backward(y, lam, rho, h if h is not None else torch.empty(0, device=y.device), x)
        return x

    @staticmethod
    def backward(ctx, xbar):
        lib = _lib.load()
        y, lam, rho, h, x = ctx.saved_tensors
        d = ctx.desc
        xbar = xbar.contiguous()
        _, _, bwd_b = lib.workspace_bytes(d)
        ws = _alloc(bwd_b, y.device)
        ybar = torch.empty_like(y)
        hbar = torch.empty_like(h) if ctx.has_h else None
        lbar, rbar = torch.empty_like(lam), torch.empty_like(rho)
        bbar = torch.empty(ctx.groups, dtype=torch.float32, device=y.device) if ctx.has_bias else None
        stream = torch.cuda.current_stream(y.device).cuda_stream
        lib.backward(d, _ptr(xbar), _ptr(x), _ptr(y), _ptr(h) if ctx.has_h else None, _ptr(lam), _ptr(rho), _ptr(ctx.ck),
                     _ptr(ybar), _ptr(hbar), _ptr(lbar), _ptr(rbar), _ptr(bbar), _ptr(ws), stream)
        return (ybar, lbar, rbar, hbar, bbar) + (None,) * 9


def tvd_fft_grouped(y, lam, rho, h=None, isotropic=False, maxit=100, *, groups: int, shared_input=False,
                    channel_concat=False, bias=None, activation="identity", creg=0.0, clamp=False, inputs_owned=False):
    """EXTENSION: ``groups`` independent reference calls of identical shape batched
    into one launch sequence (differentiable: per-group cotangents).

    * per-image PSFs / noise levels (BASELINE configs[4]): ``groups = B``; ``lam``, ``rho`` hold one value per
      image, ``h`` is ``(groups, 1, kw, kh)``; result == the reference run once per image with B = 1.
    * the parallel branches of ``net_build.jl:113-128`` (5 x ``ADMMDeconvF2((), 50, rho_i, relu1)`` on the same
      input, concatenated on channels): ``shared_input=True, channel_concat=True``; ``y`` is ``(B,P,N,M)``, the
      result ``(B, groups*P, N, M)`` -- y is read once and the result is written straight into the
      concatenated layout (no 5x re-read, no ``cat`` copy).

    Without ``shared_input`` y is ``(groups*Bg, P, N, M)`` with group g owning images ``[g*Bg, (g+1)*Bg)``.
    """
    lib = _lib.load()
    _check_cuda_f32("y", y)
    _check_cuda_f32("lambda", lam)
    _check_cuda_f32("rho", rho)
    if lam.numel() != groups or rho.numel() != groups:
        raise ValueError("lam and rho need one value per group")
    Bin, P, N, M = y.shape
    Bg = Bin if shared_input else Bin // groups
    if not shared_input and Bin % groups:
        raise ValueError("batch not divisible by groups")
    kh, kw = (0, 0)
    if h is not None and h.numel() > 0:
        _check_cuda_f32("h", h)
        if h.shape[0] != groups:
            raise ValueError("h must be (groups, 1, kw, kh)")
        kh, kw = int(h.shape[-1]), int(h.shape[-2])
    else:
        h = None
    if bias is not None:
        _check_cuda_f32("bias", bias)
    flags = (0 if clamp else _lib.FLAG_NO_CLAMP) | (_lib.FLAG_SHARED_INPUT if shared_input else 0) | \
        (_lib.FLAG_CHANNEL_CONCAT if channel_concat else 0)
    if not inputs_owned:   # inputs_owned: the caller built lam / rho / h just for this call and reads the clamp back from them
        lam, rho, h = _own_if_nonleaf(lam, clamp), _own_if_nonleaf(rho, clamp), _own_if_nonleaf(h, clamp)
    return _AdmmGroupedFunction.apply(y, lam, rho, h, bias, int(maxit), bool(isotropic), activation, float(creg), flags,
                                      int(groups), bool(shared_input), bool(channel_concat), _needs_grad(y, lam, rho, h, bias))
