"""Seeded synthetic cases shared by the CPU and GPU parity tests."""
from __future__ import annotations

import math

import numpy as np
import torch

from oracle import admm_tv_oracle as O

DT = torch.float64


def make_case(M, N, P, B, kh, kw, seed, psf="random", noise=0.02):
    """Returns (y (M,N,P,B) fp64, h (kh,kw,1,1) fp64 or None, g ground truth)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = O.synthetic_truth(M, N, P, B, seed)
    if kh == 0:
        h = None
    elif psf == "gauss":
        assert kh == kw
        h = O.gaussian_psf(kh, 2.0)
    elif psf == "motion":
        assert kh == kw
        h = O.motion_psf(kh, float(rng.uniform(0, math.pi)), float(rng.uniform(5, kh)))
    elif psf == "line":   # tests/admm_deconv_test.jl:19-20
        h = torch.zeros(kh, kw, 1, 1, dtype=DT)
        h[kh // 2, :, 0, 0] = 1.0 / kw
    else:
        h = torch.from_numpy(rng.random((kh, kw, 1, 1)))
        h = h / h.sum()
    y = O.synthetic_observation(g, h, noise, seed)
    return y, h, g


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm())


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * math.log10(1.0 / mse)
