"""Host-side mirror of the reference's Flux layers (src/layers/deconv_admm.jl).

Same names, same constructor arguments, same trainable sets, same call semantics:

    ADMMDeconv((kh,kw), num_it, σ; iso, init, groups, bias, creg)          :189-207  trainable weight,bias,λ,ρ (:209)
    ADMMDeconvF1((kh,kw), num_it, λ, σ; ...)   λ fixed                       :31-53    trainable weight,bias,ρ   (:55)
    ADMMDeconvF2((kh,kw), num_it, ρ, σ; ...)   ρ fixed                       :83-105   trainable weight,bias,λ   (:107)
    ADMMDeconvF3((kh,kw), num_it, λ, ρ, σ; ...) λ,ρ fixed                    :135-159  trainable weight,bias     (:161)

and the 8-positional forms ``ADMMDeconv(w, σ, b, λ, ρ, iters, iso, creg)`` (:176-186 etc.) via
``from_arrays``.  Calling a layer clamps λ, ρ to [creg, ∞) and weight to [0,1] *and keeps the
clamped values* (:216-219), runs tvd_fft (:221), adds the bias (:222) and applies σ (:224).

σ is one of "identity", "relu", "relu6", "relu1" (the activations net_build.jl uses, :8,107-119,155).
Tensors follow ops.py's convention: inputs are (B,P,N,M), the weight is (1,1,kw,kh).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Union

import torch
from torch import nn

from . import ops

_ACTS = ("identity", "relu", "relu6", "relu1")


def glorot_uniform(*dims: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Flux.glorot_uniform(dims...) : U(-s, s), s = sqrt(24 / sum(nfan(dims))) (gain 1).
    nfan: 1 dim -> (n, 1)... Flux: nfan(n) = (1, n); nfan(dims...) for conv = (prod(k)*cin, prod(k)*cout)
    where Julia dims are (k..., cin, cout)."""
    if len(dims) == 1:
        fan_in, fan_out = 1, dims[0]
    elif len(dims) == 2:
        fan_in, fan_out = dims[1], dims[0]
    else:
        k = math.prod(dims[:-2])
        fan_in, fan_out = k * dims[-2], k * dims[-1]
    s = math.sqrt(24.0 / (fan_in + fan_out))
    return (torch.rand(*dims, generator=generator) - 0.5) * 2 * s


class _AdmmBase(nn.Module):
    _trainable: Sequence[str] = ()

    def __init__(self, k, num_it: int, lam, rho, sigma: str, iso: bool, init, groups: int, bias: bool, creg: float):
        super().__init__()
        if sigma not in _ACTS:
            raise ValueError(f"σ must be one of {_ACTS}")
        k = tuple(k)
        if len(k) == 0:
            weight = torch.empty(0)                               # empty(ones(1)), :199
        else:
            kh, kw = k
            # Flux.convfilter(k, 1=>1; init, groups) -> (kh,kw,1,1); stored reversed: (1,1,kw,kh)
            w = (init or glorot_uniform)(kh, kw, 1, 1)
            weight = ops.from_julia(w)
        lam_t = glorot_uniform(1).abs() if lam is None else torch.zeros(1) + float(lam)    # :203 / :49
        rho_t = glorot_uniform(1).abs() if rho is None else torch.zeros(1) + float(rho)    # :204 / :102
        self.sigma = sigma
        self.iters = int(num_it)
        self.iso = bool(iso)
        self.creg = float(creg)
        self.nograd_repeat = False  # set True to reproduce train.jl:10 (`@nograd CUDA.repeat`)
        self.weight = nn.Parameter(weight.float(), requires_grad="weight" in self._trainable and weight.numel() > 0)
        # Flux.create_bias(weight, bias, 1): `false` or a zero 1-vector
        self.bias = nn.Parameter(torch.zeros(1), requires_grad="bias" in self._trainable) if bias else None
        self.lam = nn.Parameter(lam_t.float(), requires_grad="lam" in self._trainable)
        self.rho = nn.Parameter(rho_t.float(), requires_grad="rho" in self._trainable)

    @classmethod
    def from_arrays(cls, w, sigma, b, lam, rho, iters, iso, creg):
        """The 8-positional constructor (w, σ, b, λ, ρ, iters, iso, creg), deconv_admm.jl:176-186."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.sigma, self.iters, self.iso, self.creg, self.nograd_repeat = sigma, int(iters), bool(iso), float(creg), False
        w = torch.as_tensor(w, dtype=torch.float32)
        self.weight = nn.Parameter(w, requires_grad="weight" in cls._trainable and w.numel() > 0)
        self.bias = None if b is None or b is False else nn.Parameter(torch.as_tensor(b, dtype=torch.float32).reshape(1),
                                                                     requires_grad="bias" in cls._trainable)
        self.lam = nn.Parameter(torch.as_tensor(lam, dtype=torch.float32).reshape(1), requires_grad="lam" in cls._trainable)
        self.rho = nn.Parameter(torch.as_tensor(rho, dtype=torch.float32).reshape(1), requires_grad="rho" in cls._trainable)
        return self

    def per_iteration_(self):
        """EXTENSION (BASELINE configs[2] "learned-rho/lambda iterations"): give every unrolled iteration
        its own (λ_k, ρ_k), initialised to the current values.  The reference has one pair for all iterations
        (deconv_admm.jl:168-169, ops.jl:20); with equal entries the result is bit-identical to it.  In place; returns self."""
        for n in ("lam", "rho"):
            p = getattr(self, n)
            if p.numel() == 1:
                setattr(self, n, nn.Parameter(p.detach().repeat(self.iters).contiguous(), requires_grad=p.requires_grad))
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(d::Admm)(x), deconv_admm.jl:215-225.  λ, ρ, weight are clamped in place by the kernel
        library, which is the reference's write-back into the struct."""
        h = self.weight if self.weight.numel() > 0 else None
        return ops.admm_layer_call(x, self.lam, self.rho, h, self.bias, self.iters, self.iso, self.sigma, self.creg,
                                   self.nograd_repeat, clamp=True)

    def trainable(self):
        return tuple(n for n in self._trainable if getattr(self, n) is not None)

    def packed_grads(self) -> torch.Tensor:
        """[hbar..., lambar, rhobar, biasbar] in one buffer (the one-call NCCL allreduce payload)."""
        parts = []
        for n in ("weight", "lam", "rho", "bias"):
            p = getattr(self, n)
            if p is not None and p.requires_grad and p.numel() > 0:
                parts.append((p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1))
        return torch.cat(parts) if parts else torch.empty(0)

    def unpack_grads(self, buf: torch.Tensor):
        o = 0
        for n in ("weight", "lam", "rho", "bias"):
            p = getattr(self, n)
            if p is not None and p.requires_grad and p.numel() > 0:
                p.grad = buf[o:o + p.numel()].reshape(p.shape).clone()
                o += p.numel()


def _check_pos(name, v):
    assert v > 0, f"Parameter {name} must be greater than 0"   # :42,94,147-148


class ADMMDeconv(_AdmmBase):
    _trainable = ("weight", "bias", "lam", "rho")               # :209

    def __init__(self, k, num_it, sigma="identity", *, iso=False, init=None, groups=1, bias=False, creg=0.0):
        super().__init__(k, num_it, None, None, sigma, iso, init, groups, bias, creg)


class ADMMDeconvF1(_AdmmBase):
    _trainable = ("weight", "bias", "rho")                      # :55

    def __init__(self, k, num_it, lam, sigma="identity", *, iso=False, init=None, groups=1, bias=False, creg=0.0):
        _check_pos("λ", lam)
        super().__init__(k, num_it, lam, None, sigma, iso, init, groups, bias, creg)


class ADMMDeconvF2(_AdmmBase):
    _trainable = ("weight", "bias", "lam")                      # :107

    def __init__(self, k, num_it, rho, sigma="identity", *, iso=False, init=None, groups=1, bias=False, creg=0.0):
        _check_pos("ρ", rho)
        super().__init__(k, num_it, None, rho, sigma, iso, init, groups, bias, creg)


class ADMMDeconvF3(_AdmmBase):
    _trainable = ("weight", "bias")                             # :161

    def __init__(self, k, num_it, lam, rho, sigma="identity", *, iso=False, init=None, groups=1, bias=False, creg=0.0):
        _check_pos("λ", lam)
        _check_pos("ρ", rho)
        super().__init__(k, num_it, lam, rho, sigma, iso, init, groups, bias, creg)


Admm = Union[ADMMDeconv, ADMMDeconvF1, ADMMDeconvF2, ADMMDeconvF3]   # :212


class ADMMParallel(nn.Module):
    """``Parallel(chcat, deconv_1, ..., deconv_G)`` of net_build.jl:113-128 (get_denoiser) as ONE grouped call:
    the G ADMM layers read the same input, their results are concatenated on the channel dimension.
    Requires layers of identical iters / iso / σ / creg, the same PSF size (or all without a PSF) and no bias
    mix; falls back to nothing -- incompatible layers raise."""

    def __init__(self, *layers: _AdmmBase):
        super().__init__()
        if len(layers) < 1:
            raise ValueError("need at least one layer")
        l0 = layers[0]
        for l in layers:
            same = (l.iters, l.iso, l.sigma, l.creg, tuple(l.weight.shape), l.bias is None) == \
                (l0.iters, l0.iso, l0.sigma, l0.creg, tuple(l0.weight.shape), l0.bias is None)
            if not same:
                raise ValueError("ADMMParallel needs layers of identical shape / iterations / iso / activation")
        self.layers = nn.ModuleList(layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        L = list(self.layers)
        G = len(L)
        lam = torch.cat([l.lam for l in L])
        rho = torch.cat([l.rho for l in L])
        h = torch.cat([l.weight for l in L], dim=0).contiguous() if L[0].weight.numel() > 0 else None
        bias = torch.cat([l.bias for l in L]) if L[0].bias is not None else None
        out = ops.tvd_fft_grouped(x, lam, rho, h, L[0].iso, L[0].iters, groups=G, shared_input=True, channel_concat=True,
                                  bias=bias, activation=L[0].sigma, creg=L[0].creg, clamp=True, inputs_owned=True)
        # the library clamped the packed copies in place: persist into the layers (deconv_admm.jl:216-219)
        with torch.no_grad():
            for g, l in enumerate(L):
                l.lam.copy_(lam[g:g + 1]); l.rho.copy_(rho[g:g + 1])
                if h is not None:
                    l.weight.copy_(h[g:g + 1])
        return out
