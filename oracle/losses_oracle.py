"""CPU oracle for the losses adjacent to the ADMM-TV path  --  TEST INFRASTRUCTURE ONLY.

SURVEY.md section 8 row f-2: the two losses that produce the cotangent the layer's pullback consumes,
``gmsd_loss`` (the loss of train.jl:191) and ``ssim_loss`` (the loss of train_v2.jl:89).  Only ``tests/``
and the benchmark tools' checking legs may import this file; the product package never does.

PARITY UNPINNED (same situation as ``admm_tv_oracle.py``): no Julia toolchain here, the reference has no
test for either loss, and ``conv`` / ``pad_circular`` live in NNlib 0.9.21 (un-vendored).  This is a
line-by-line restatement of

    src/metrics/iqa_utils.jl:15-20   SOBEL_KERNEL_X / SOBEL_KERNEL_Y
    src/metrics/iqa_utils.jl:24-50   imgrads
    src/metrics/iqa_utils.jl:53-55   gradientsmag
    src/metrics/gmsd.jl:5-10         similarity_map
    src/metrics/gmsd.jl:13-27        gmsd            (:30 gmsd_loss = gmsd)
    src/metrics/ssim.jl:6-17         SSIM_KERNEL
    src/metrics/ssim.jl:25-47        ssim_kernel
    src/metrics/ssim.jl:84-124       ssim            (:148 ssim_loss = 1 - ssim, :160-164 ssim_loss_fast),
                                     incl. crop=false (:104-110, NNlib pad_symmetric) and arbitrary windows

pinned by self-consistency only (tests/test_losses_oracle.py): gmsd(x,x) = 0, ssim(x,x) = 1, symmetry,
shift invariance of gmsd (circular padding), an independent roll-based formulation, finite differences
of the gradients.  Arrays are indexed like the Julia arrays, ``a[i, j, c, b]`` of shape (M,N,C,B);
``torch.autograd`` through the fp64 restatement stands in for Zygote.
"""
from __future__ import annotations

import torch

from .admm_tv_oracle import nnlib_conv, pad_circular

# src/metrics/ssim.jl:6-17
SSIM_KERNEL = [0.00102838008447911, 0.007598758135239185, 0.03600077212843083, 0.10936068950970002,
               0.2130055377112537, 0.26601172486179436, 0.2130055377112537, 0.10936068950970002,
               0.03600077212843083, 0.007598758135239185, 0.00102838008447911]


def sobel_kernels(dtype=torch.float64):
    """iqa_utils.jl:15-20.  ``cat(c1, c2, c3, dims=2)`` builds a matrix whose COLUMNS are the vectors."""
    kx = torch.tensor([[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]], dtype=dtype) / 8.0
    return kx, kx.t().contiguous()


def imgrads(x: torch.Tensor):
    """iqa_utils.jl:24-50: grouped (per-channel) true convolution of the circularly padded image."""
    C = x.shape[2]
    kx, ky = sobel_kernels(x.dtype)
    wx = kx.reshape(3, 3, 1, 1).repeat(1, 1, 1, C)      # :37-38 repeat(ker, 1,1,1,groups)
    wy = ky.reshape(3, 3, 1, 1).repeat(1, 1, 1, C)
    xp = pad_circular(x, (1, 1, 1, 1))                   # :44 padding = (3-1)/2
    return nnlib_conv(xp, wx, groups=C), nnlib_conv(xp, wy, groups=C)   # :46-47


def gradientsmag(gx, gy):
    """iqa_utils.jl:53-55"""
    return torch.sqrt(gx ** 2 + gy ** 2 + 1e-16)


def similarity_map(map_xref, map_x, constant, alpha):
    """gmsd.jl:5-10"""
    num = 2.0 * map_xref * map_x - alpha * map_xref * map_x + constant
    den = map_xref ** 2 + map_x ** 2 - alpha * map_xref * map_x + constant
    return num / den


def gmsd(x: torch.Tensor, y: torch.Tensor, t: float = 0.0026, alpha: float = 0.0) -> torch.Tensor:
    """gmsd.jl:13-27 with reduction = mean."""
    xgx, xgy = imgrads(x)
    ygx, ygy = imgrads(y)
    map_x = gradientsmag(xgx, xgy)
    map_y = gradientsmag(ygx, ygy)
    gms = similarity_map(map_x, map_y, t, alpha)
    mean_gms = gms.mean(dim=(0, 1, 2), keepdim=True)
    score = ((gms - mean_gms) ** 2).mean(dim=(0, 1, 2), keepdim=True)
    return torch.sqrt(score).mean()


gmsd_loss = gmsd   # gmsd.jl:30


def gmsd_roll(x: torch.Tensor, y: torch.Tensor, t: float = 0.0026, alpha: float = 0.0) -> torch.Tensor:
    """Independent formulation (torch.roll stencils) used to cross-check the literal restatement."""
    def grads(a):
        r = lambda d1, d2: torch.roll(a, shifts=(-d1, -d2), dims=(0, 1))   # r(d1,d2)[i,j] = a[i+d1, j+d2]
        gx = sum(w * (r(1, dj) - r(-1, dj)) for dj, w in ((-1, 1.0), (0, 2.0), (1, 1.0))) / 8.0
        gy = sum(w * (r(di, 1) - r(di, -1)) for di, w in ((-1, 1.0), (0, 2.0), (1, 1.0))) / 8.0
        return gx, gy
    mx = gradientsmag(*grads(x))
    my = gradientsmag(*grads(y))
    g = similarity_map(mx, my, t, alpha)
    m = g.mean(dim=(0, 1, 2), keepdim=True)
    return torch.sqrt(((g - m) ** 2).mean(dim=(0, 1, 2))).mean()


def ssim_kernel(dtype=torch.float64, length=None):
    """ssim.jl:25-47 for 4-D inputs (N-2 == 2): outer product of the 11-tap Gaussian, (11,11,1,1);
    ``length``: the normalised box kernel of ssim_loss_fast (ssim.jl:160-164) instead."""
    if length is not None:
        k = torch.ones(length, length, dtype=dtype)
        return (k / k.sum()).reshape(length, length, 1, 1)
    g = torch.tensor(SSIM_KERNEL, dtype=dtype)
    return torch.outer(g, g).reshape(11, 11, 1, 1)


def pad_symmetric(x: torch.Tensor, pads) -> torch.Tensor:
    """NNlib.pad_symmetric(x, (d1_lo, d1_hi, d2_lo, d2_hi)) on an (M,N,C,B) array: the values are mirrored across the
    border INCLUDING the border sample ([b a | a b c d | d c] for pads (2, 2)), unlike pad_reflect."""
    lo1, hi1, lo2, hi2 = pads
    M, N = x.shape[0], x.shape[1]

    def mirror(lo, hi, n):
        idx = []
        for t in range(-lo, n + hi):
            idx.append(-1 - t if t < 0 else (2 * n - 1 - t if t >= n else t))
        return torch.tensor(idx, dtype=torch.long)
    return x.index_select(0, mirror(lo1, hi1, M)).index_select(1, mirror(lo2, hi2, N))


def ssim(x: torch.Tensor, y: torch.Tensor, kernel: torch.Tensor | None = None, peakval: float = 1.0,
         crop: bool = True) -> torch.Tensor:
    """ssim.jl:84-124 (``dims`` is never read by the reference's body).  ``kernel``: (L1, L2, 1, C or 1)."""
    C = x.shape[2]
    kernel = ssim_kernel(x.dtype) if kernel is None else kernel
    if kernel.shape[3] != C:
        kernel = kernel.repeat(1, 1, 1, C)               # :96-98
    C1, C2 = (peakval * 0.01) ** 2, (peakval * 0.03) ** 2  # :101-102
    if not crop:                                          # :104-110  calc_padding of Flux's conv.jl
        k1, k2 = kernel.shape[0] - 1, kernel.shape[1] - 1
        padding = (-(-k1 // 2), k1 // 2, -(-k2 // 2), k2 // 2)   # (cld, fld) per dimension
        x = pad_symmetric(x, padding)
        y = pad_symmetric(y, padding)
    conv = lambda a: nnlib_conv(a, kernel, groups=C)
    mx, my = conv(x), conv(y)                             # :112-113
    mx2, my2, mxy = mx ** 2, my ** 2, mx * my
    sx2 = conv(x ** 2) - mx2                              # :117
    sy2 = conv(y ** 2) - my2
    sxy = conv(x * y) - mxy
    smap = (2 * mxy + C1) * (2 * sxy + C2) / ((mx2 + my2 + C1) * (sx2 + sy2 + C2))   # :121
    return smap.mean(dim=(0, 1, 2)).mean()                # :122-123


def ssim_loss(x, y, kernel=None, peakval: float = 1.0, crop: bool = True):
    """ssim.jl:148"""
    return 1.0 - ssim(x, y, kernel, peakval, crop)


def ssim_loss_fast(x, y, kernel_length: int = 5, peakval: float = 1.0, crop: bool = True):
    """ssim.jl:160-164"""
    return ssim_loss(x, y, ssim_kernel(x.dtype, kernel_length), peakval, crop)
