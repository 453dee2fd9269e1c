"""CPU oracle for the ADMM-TV deconvolution layer  --  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package (``admm_deconv_b200``) never imports anything under ``oracle/``.

PARITY UNPINNED.  The reference (georgegrosu1/admm-deconv) is Julia + FFTW + NNlib + Zygote.
There is no ``julia`` binary in this image, the reference's only test
(``src/tests/admm_deconv_test.jl``) holds no assertion and no stored value, and the arithmetic
partly lives in third-party packages that are not vendored (NNlib 0.9.21 ``conv`` /
``pad_circular`` / ``pad_constant``, FFTW.jl 1.8.0 ``rfft`` / ``irfft``, Zygote 0.6.70 for the
backward; pins in ``Manifest.toml``).  This oracle is therefore a *restatement*, line by line,
of

    src/ops/ops.jl:17-96      tvd_fft_cpu        (the GPU twin :99-178 is the same arithmetic)
    src/ops/ops.jl:6,9,10     pixelnorm, ST, BT
    src/layers/deconv_admm.jl:215-225   the layer call (clamp -> tvd_fft -> +bias -> sigma)

with NNlib's published conventions restated in ``pad_circular`` / ``pad_constant`` /
``nnlib_conv`` below (``conv`` is a true convolution, i.e. the kernel is flipped;
``flipkernel=false`` default).  That convention is from the NNlib documentation and cannot be
executed here; it is kept behind the single switch ``NNLIB_CONV_FLIPS_KERNEL``.
It is pinned only by self-consistency (tests/test_oracle.py): adjointness of D/D^T and H/H^T,
known answers (delta PSF, constant image, K=1 closed form), an asymmetric-PSF restoration
check that fails if the flip convention is wrong, and agreement between this literal
restatement and an independent roll/spectral formulation (``tvd_fft_fast``).

Array convention: every array is indexed exactly like the Julia array, ``a[i, j, p, b]`` with
shape ``(M, N, P, B)`` (0-based here).  Memory order is irrelevant to the oracle; the product's
C ABI uses Julia's column-major order (dim 1 contiguous), and the ctypes harness converts.

Gradients: ``torch.autograd`` through the fp64 restatement stands in for Zygote.  Both
differentiate the same primal program, so the cotangents are the same mathematical objects
(SURVEY.md section 8a-10).
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F

# NNlib.conv(x, w) with the default flipkernel=false computes a true convolution (the kernel
# is flipped relative to cross-correlation).  Flux ships CrossCor for the unflipped case.
NNLIB_CONV_FLIPS_KERNEL = True


# --------------------------------------------------------------------------------------------
# NNlib primitives restated (call sites: ops.jl:25,62-65,78-81)
# --------------------------------------------------------------------------------------------
def pad_circular(x: torch.Tensor, pads: Tuple[int, int, int, int]) -> torch.Tensor:
    """NNlib.pad_circular(x, (d1_lo, d1_hi, d2_lo, d2_hi)) on an (M,N,C,B) array."""
    lo1, hi1, lo2, hi2 = pads
    M, N = x.shape[0], x.shape[1]
    idx1 = (torch.arange(-lo1, M + hi1) % M).to(torch.long)
    idx2 = (torch.arange(-lo2, N + hi2) % N).to(torch.long)
    return x.index_select(0, idx1).index_select(1, idx2)


def pad_constant(x: torch.Tensor, pads, value: float = 0.0) -> torch.Tensor:
    """NNlib.pad_constant(x, (d1_lo,d1_hi,d2_lo,d2_hi,d3_lo,d3_hi,d4_lo,d4_hi))."""
    pads = tuple(pads) + (0,) * (8 - len(pads))
    # torch pads from the LAST dim backwards
    tp = (pads[6], pads[7], pads[4], pads[5], pads[2], pads[3], pads[0], pads[1])
    return F.pad(x, tp, mode="constant", value=value)


def nnlib_conv(x: torch.Tensor, w: torch.Tensor, groups: int = 1) -> torch.Tensor:
    """NNlib.conv(x, w, DenseConvDims(x, w; groups)) : stride 1, no padding.

    x is (H, W, Cin, Bt); w is (kh, kw, Cin/groups, Cout); result (H-kh+1, W-kw+1, Cout, Bt).
    out[i,j,co,b] = sum_{a,b',ci} w[a,b',ci,co] * x[i+kh-1-a, j+kw-1-b', g(co)*cpg+ci, b]   (0-based)
    """
    xt = x.permute(3, 2, 0, 1)  # (Bt, Cin, H, W)
    wt = w.permute(3, 2, 0, 1)  # (Cout, Cin/groups, kh, kw)
    if NNLIB_CONV_FLIPS_KERNEL:
        wt = torch.flip(wt, dims=(2, 3))
    out = F.conv2d(xt, wt, groups=groups)
    return out.permute(2, 3, 1, 0)


def rfft12(x: torch.Tensor) -> torch.Tensor:
    """Julia rfft(x, (1,2)) (or rfft(x) on a matrix): dim 1 is the halved dimension."""
    return torch.fft.rfftn(x, dim=(1, 0))


def irfft12(X: torch.Tensor, M: int) -> torch.Tensor:
    """Julia irfft(X, M, (1,2)) (normalised)."""
    N = X.shape[1]
    return torch.fft.irfftn(X, s=(N, M), dim=(1, 0))


# --------------------------------------------------------------------------------------------
# prox operators  (ops.jl:6,9,10)
# --------------------------------------------------------------------------------------------
def pixelnorm(x: torch.Tensor) -> torch.Tensor:
    """ops.jl:6   sqrt.(sum(x.^2, dims=(3,4)))  -- ONE norm per pixel over dims 3 and 4."""
    return torch.sqrt(torch.sum(x * x, dim=(2, 3), keepdim=True))


def ST(x: torch.Tensor, tau: torch.Tensor) -> torch.Tensor:
    """ops.jl:9   sign.(x).*max.(abs.(x).-tau, 0f0)"""
    return torch.sign(x) * torch.clamp(torch.abs(x) - tau, min=0.0)


def BT(x: torch.Tensor, tau: torch.Tensor) -> torch.Tensor:
    """ops.jl:10  max.(1 .- tau ./ pixelnorm(x), 0).*x  (n = 0 gives max(-Inf,0)*0 = 0)."""
    return torch.clamp(1.0 - tau / pixelnorm(x), min=0.0) * x


# --------------------------------------------------------------------------------------------
# the solver, literal  (ops.jl:17-96)
# --------------------------------------------------------------------------------------------
def tvd_fft_cpu(
    y: torch.Tensor,
    lam: torch.Tensor,
    rho: torch.Tensor,
    h: Optional[torch.Tensor] = None,
    isotropic: bool = False,
    maxit: int = 100,
    *,
    nograd_repeat: bool = False,
    trace: Optional[list] = None,
) -> torch.Tensor:
    """Line-by-line restatement of tvd_fft_cpu (ops.jl:17-96).

    y (M,N,P,B); lam, rho 1-element tensors; h (kh,kw,1,1) or None / empty.
    ``nograd_repeat`` reproduces train.jl:10 (``Zygote.@nograd CUDA.repeat``): the gradient
    through ``h = repeat(h,1,1,1,B)`` (ops.jl:71), i.e. through the spatial H^T(y) path, is cut.
    ``trace``: if a list, (x, z, u) of every iteration is appended (test use).
    """
    T = y.dtype
    M, N, P, B = y.shape                                   # ops.jl:18
    y = y.permute(0, 1, 3, 2)                              # ops.jl:19  (M,N,B,P)
    tau = lam / rho                                        # ops.jl:20

    h_empty = h is None or h.numel() == 0
    if h_empty:                                            # ops.jl:22-23
        Sigma = torch.ones(1, 1, 1, 1, dtype=T)
    else:                                                  # ops.jl:25-27
        hh = pad_constant(h, (0, M - h.shape[0], 0, N - h.shape[1], 0, 0, 0, 0))
        Sigma_ref = rfft12(hh[:, :, 0, 0])
        Sigma = Sigma_ref.reshape(Sigma_ref.shape[0], Sigma_ref.shape[1], 1, 1)

    # ops.jl:32-37
    dx_filter = torch.zeros(M, N, dtype=T)
    dx_filter[0, 0] = 1.0
    dx_filter[0, 1] = -1.0
    dy_filter = torch.zeros(M, N, dtype=T)
    dy_filter[0, 0] = 1.0
    dy_filter[1, 0] = -1.0
    Lx = rfft12(dx_filter)
    Ly = rfft12(dy_filter)
    Lsum = (Lx.abs() ** 2 + Ly.abs() ** 2).reshape(M // 2 + 1, N, 1, 1)
    C = 1.0 / (Sigma.abs() ** 2 + rho * Lsum)

    thresh_type: Callable = BT if isotropic else ST        # ops.jl:39-43

    x = torch.zeros(M, N, B, P, dtype=T)                   # ops.jl:46-49
    z = torch.zeros(M, N, 2 * B, P, dtype=T)
    u = torch.zeros(M, N, 2 * B, P, dtype=T)

    # ops.jl:52-59  (W is (2,2,1,2B); W^T is (2,2,2,B))
    W1 = torch.tensor([[1.0, -1.0], [0.0, 0.0]], dtype=T)
    W2 = torch.tensor([[1.0, 0.0], [-1.0, 0.0]], dtype=T)
    W = torch.stack([W1, W2], dim=-1).reshape(2, 2, 1, 2).repeat(1, 1, 1, B)
    Wt1 = torch.tensor([[0.0, 0.0], [-1.0, 1.0]], dtype=T)
    Wt2 = torch.tensor([[0.0, -1.0], [0.0, 1.0]], dtype=T)
    Wt = torch.stack([Wt1, Wt2], dim=-1).reshape(2, 2, 1, 2).permute(0, 1, 3, 2).repeat(1, 1, 1, B)

    def D(x_):                                             # ops.jl:64
        return nnlib_conv(pad_circular(x_, (1, 0, 1, 0)), W, groups=B)

    def Dt(z_):                                            # ops.jl:65
        return nnlib_conv(pad_circular(z_, (0, 1, 0, 1)), Wt, groups=B)

    if h_empty:                                            # ops.jl:67-69
        Ht = lambda a: a
    else:                                                  # ops.jl:71-81
        hr = (h.detach() if nograd_repeat else h).repeat(1, 1, 1, B)
        ht = torch.flip(hr, dims=(0, 1, 2, 3))
        kh, kw = h.shape[0], h.shape[1]
        padu, padd = math.ceil((kh - 1) / 2), math.floor((kh - 1) / 2)
        padl, padr = math.ceil((kw - 1) / 2), math.floor((kw - 1) / 2)
        pad2 = (padd, padu, padr, padl)
        Ht = lambda a: nnlib_conv(pad_circular(a, pad2), ht, groups=B)

    for _ in range(maxit):                                 # ops.jl:84-92
        x = irfft12(C * rfft12(Ht(y) + rho * Dt(z - u)), M)
        Dxk = D(x)
        z = thresh_type(Dxk + u, tau)
        u = u + Dxk - z
        if trace is not None:
            trace.append((x.detach().clone(), z.detach().clone(), u.detach().clone()))
    return x.permute(0, 1, 3, 2)                           # ops.jl:93


def H_forward(x: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """The reference's blur operator H (ops.jl:73-80) on an (M,N,P,B) array; used to make
    synthetic observations y = H g + noise with the reference's own alignment."""
    M, N, P, B = x.shape
    xg = x.permute(0, 1, 3, 2)
    kh, kw = h.shape[0], h.shape[1]
    padu, padd = math.ceil((kh - 1) / 2), math.floor((kh - 1) / 2)
    padl, padr = math.ceil((kw - 1) / 2), math.floor((kw - 1) / 2)
    hr = h.repeat(1, 1, 1, B)
    out = nnlib_conv(pad_circular(xg, (padu, padd, padl, padr)), hr, groups=B)
    return out.permute(0, 1, 3, 2)


# --------------------------------------------------------------------------------------------
# the layer call  (deconv_admm.jl:215-225)
# --------------------------------------------------------------------------------------------
ACTIVATIONS = {
    "identity": lambda t: t,
    "relu": lambda t: torch.clamp(t, min=0.0),
    "relu6": lambda t: torch.clamp(t, min=0.0, max=6.0),
    "relu1": lambda t: torch.clamp(t, min=0.0, max=1.0),   # net_build.jl:8
}


def admm_layer(
    x: torch.Tensor,
    weight: Optional[torch.Tensor],
    bias,
    lam: torch.Tensor,
    rho: torch.Tensor,
    iters: int,
    iso: bool = False,
    creg: float = 0.0,
    act: str = "identity",
    *,
    nograd_repeat: bool = False,
):
    """(d::Admm)(x), deconv_admm.jl:215-225.  Returns (out, (lam_c, rho_c, weight_c)): the clamped
    parameters are what the reference writes back into the struct (:216-219)."""
    lam_c = torch.clamp(lam, min=creg)                     # :216
    rho_c = torch.clamp(rho, min=creg)                     # :217
    w_c = None if weight is None or weight.numel() == 0 else torch.clamp(weight, 0.0, 1.0)  # :219
    res = tvd_fft_cpu(x, lam_c, rho_c, w_c, iso, iters, nograd_repeat=nograd_repeat)        # :221
    if bias is not None and bias is not False:
        res = res + bias                                   # :222
    return ACTIVATIONS[act](res), (lam_c, rho_c, w_c)      # :224


# --------------------------------------------------------------------------------------------
# independent formulation (roll differences, spectral / hoisted H^T y) -- cross-check + speed
# --------------------------------------------------------------------------------------------
def D_roll(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dim-2 difference, dim-1 difference) = channels (2s-1, 2s) of D(x); SURVEY 8a-7."""
    return x - torch.roll(x, 1, dims=1), x - torch.roll(x, 1, dims=0)


def Dt_roll(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """D^T; SURVEY 8a-6."""
    return (t1 - torch.roll(t1, -1, dims=1)) + (t2 - torch.roll(t2, -1, dims=0))


def Ht_roll(y: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """H^T y as an explicit shifted sum: sum_{a,b} h[a,b] y[i+a-pd, j+b-pr] (0-based a,b)."""
    kh, kw = h.shape[0], h.shape[1]
    pd, pr = (kh - 1) // 2, (kw - 1) // 2
    out = torch.zeros_like(y)
    for a in range(kh):
        for b in range(kw):
            out = out + h[a, b, 0, 0] * torch.roll(y, shifts=(-(a - pd), -(b - pr)), dims=(0, 1))
    return out


def spectral_tables(M: int, N: int, h: Optional[torch.Tensor], rho: torch.Tensor, dtype):
    """Sigma (rfft of the corner-placed PSF) and C on the (M/2+1, N) half spectrum, analytic
    |Lambda|^2 = 4 sin^2(pi k/N): SURVEY 8a-5."""
    k1 = torch.arange(M // 2 + 1, dtype=dtype).reshape(-1, 1)
    k2 = torch.arange(N, dtype=dtype).reshape(1, -1)
    L = 4 * torch.sin(math.pi * k2 / N) ** 2 + 4 * torch.sin(math.pi * k1 / M) ** 2
    if h is None or h.numel() == 0:
        Sig = torch.ones(M // 2 + 1, N, dtype=torch.complex128 if dtype == torch.float64 else torch.complex64)
    else:
        hh = torch.zeros(M, N, dtype=dtype)
        hh[: h.shape[0], : h.shape[1]] = h[:, :, 0, 0]
        Sig = rfft12(hh)
    C = 1.0 / (Sig.abs() ** 2 + rho * L)
    return Sig, C


def tvd_fft_fast(
    y: torch.Tensor,
    lam: torch.Tensor,
    rho: torch.Tensor,
    h: Optional[torch.Tensor] = None,
    isotropic: bool = False,
    maxit: int = 100,
    hoist: bool = True,
    states: Optional[list] = None,
) -> torch.Tensor:
    """Same mathematics as tvd_fft_cpu, written independently: roll-based D / D^T, analytic
    |Lambda|^2, explicit shifted-sum H^T y (hoisted out of the loop when ``hoist``).
    Works directly on (M,N,P,B) (no permute; the permute only regroups independent planes).

    EXTENSION (not in the reference): when ``lam`` / ``rho`` hold ``maxit`` values, iteration k uses
    (lam[k-1], rho[k-1]) for its x-update and its shrinkage -- "learned per-iteration parameters".  With one value each
    this is exactly the reference recursion (ops.jl:84-92)."""
    M, N, P, B = y.shape
    per_iter = lam.numel() > 1 or rho.numel() > 1
    if per_iter:
        assert lam.numel() == maxit and rho.numel() == maxit, "per-iteration parameters: one value per iteration"
    lam_k = lambda k: lam[k:k + 1] if per_iter else lam
    rho_k = lambda k: rho[k:k + 1] if per_iter else rho
    C_all = [spectral_tables(M, N, h, rho_k(k), y.dtype)[1].reshape(M // 2 + 1, N, 1, 1) for k in range(maxit if per_iter else 1)]
    h_empty = h is None or h.numel() == 0
    Hty = (lambda: y) if h_empty else (lambda: Ht_roll(y, h))
    b = Hty() if hoist else None
    z1 = torch.zeros_like(y); z2 = torch.zeros_like(y)
    u1 = torch.zeros_like(y); u2 = torch.zeros_like(y)
    x = torch.zeros_like(y)
    for k in range(maxit):
        bb = b if hoist else Hty()
        rho_, tau, C = rho_k(k), lam_k(k) / rho_k(k), C_all[k if per_iter else 0]
        x = irfft12(C * rfft12(bb + rho_ * Dt_roll(z1 - u1, z2 - u2)), M)
        d1, d2 = D_roll(x)
        v1, v2 = d1 + u1, d2 + u2
        if isotropic:
            n = torch.sqrt(torch.sum(v1 * v1 + v2 * v2, dim=(2, 3), keepdim=True))
            s = torch.clamp(1.0 - tau / n, min=0.0)
            z1, z2 = s * v1, s * v2
        else:
            z1, z2 = ST(v1, tau), ST(v2, tau)
        u1, u2 = (u1 + d1) - z1, (u2 + d2) - z2
        if states is not None:
            states.append((x.clone(), v1.clone(), v2.clone()))
    return x


# --------------------------------------------------------------------------------------------
# reference gradients via autograd (stands in for Zygote)
# --------------------------------------------------------------------------------------------
def layer_grads(
    x: torch.Tensor,
    xbar: torch.Tensor,
    weight: Optional[torch.Tensor],
    bias,
    lam: torch.Tensor,
    rho: torch.Tensor,
    iters: int,
    iso: bool = False,
    creg: float = 0.0,
    act: str = "identity",
    nograd_repeat: bool = False,
):
    """Returns (out, dict of cotangents) for the layer call given the output cotangent xbar."""
    x = x.detach().clone().requires_grad_(True)
    lam = lam.detach().clone().requires_grad_(True)
    rho = rho.detach().clone().requires_grad_(True)
    has_w = weight is not None and weight.numel() > 0
    w = weight.detach().clone().requires_grad_(True) if has_w else None
    has_b = bias is not None and bias is not False
    bvar = bias.detach().clone().requires_grad_(True) if has_b else None
    out, _ = admm_layer(x, w, bvar, lam, rho, iters, iso, creg, act, nograd_repeat=nograd_repeat)
    wrt = [x, lam, rho] + ([w] if has_w else []) + ([bvar] if has_b else [])
    g = torch.autograd.grad(out, wrt, grad_outputs=xbar, allow_unused=True)
    res = {"x": g[0], "lam": g[1], "rho": g[2]}
    k = 3
    if has_w:
        res["weight"] = g[k]; k += 1
    if has_b:
        res["bias"] = g[k]
    return out.detach(), res


# --------------------------------------------------------------------------------------------
# synthetic data  
# --------------------------------------------------------------------------------------------
def gaussian_psf(k: int, sigma: float, dtype=torch.float64) -> torch.Tensor:
    ax = torch.arange(k, dtype=torch.float64) - (k - 1) / 2
    g = torch.exp(-(ax ** 2) / (2 * sigma ** 2))
    p = torch.outer(g, g)
    return (p / p.sum()).to(dtype).reshape(k, k, 1, 1)


def motion_psf(k: int, theta: float, length: float, dtype=torch.float64) -> torch.Tensor:
    """Linear motion blur through the centre of a k x k support, bilinear rasterised."""
    p = torch.zeros(k, k, dtype=torch.float64)
    c = (k - 1) / 2
    n = max(int(math.ceil(length * 4)), 2)
    for s in torch.linspace(-length / 2, length / 2, n).tolist():
        a, b = c + s * math.sin(theta), c + s * math.cos(theta)
        a0, b0 = int(math.floor(a)), int(math.floor(b))
        for da in (0, 1):
            for db in (0, 1):
                aa, bb = a0 + da, b0 + db
                if 0 <= aa < k and 0 <= bb < k:
                    p[aa, bb] += (1 - abs(a - aa)) * (1 - abs(b - bb))
    return (p / p.sum()).to(dtype).reshape(k, k, 1, 1)


def synthetic_truth(M: int, N: int, P: int, B: int, seed: int, dtype=torch.float64) -> torch.Tensor:
    """Box-filtered uniform noise + 8 random rectangles per plane, min-max to [0,1]."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.random((M, N, P, B))
    t = torch.from_numpy(g)
    ker = 9
    acc = torch.zeros_like(t)
    for a in range(-(ker // 2), ker // 2 + 1):
        acc = acc + torch.roll(t, a, dims=0)
    acc2 = torch.zeros_like(t)
    for a in range(-(ker // 2), ker // 2 + 1):
        acc2 = acc2 + torch.roll(acc, a, dims=1)
    t = acc2
    mn = t.amin(dim=(0, 1), keepdim=True); mx = t.amax(dim=(0, 1), keepdim=True)
    t = (t - mn) / (mx - mn)
    for b in range(B):
        for p in range(P):
            for _ in range(8):
                i0, i1 = sorted(rng.integers(0, M, 2).tolist()); j0, j1 = sorted(rng.integers(0, N, 2).tolist())
                t[i0:i1 + 1, j0:j1 + 1, p, b] = float(rng.random())
    return t.to(dtype)


def synthetic_observation(g: torch.Tensor, h: Optional[torch.Tensor], noise_sigma: float, seed: int) -> torch.Tensor:
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    y = g if h is None or h.numel() == 0 else H_forward(g, h.to(g.dtype))
    return y + noise_sigma * torch.from_numpy(rng.standard_normal(tuple(g.shape))).to(g.dtype)
