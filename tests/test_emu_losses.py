"""CPU tests of the loss kernels (include/admmtv_loss.h) through the EMULATED build of the kernel sources
against the fp64 oracle restatement of gmsd.jl / ssim.jl (value and gradient w.r.t. the prediction)."""
import numpy as np
import pytest
import torch

import emu_harness as E
from oracle import losses_oracle as LO

DT = torch.float64


def _images(M, N, C, B, seed, noise=0.08):
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.random((M, N, C, B))
    # smooth a little so gradient magnitudes look like images, then perturb for the prediction
    g = (g + np.roll(g, 1, 0) + np.roll(g, 1, 1) + np.roll(g, (1, 1), (0, 1))) / 4
    x = g + noise * rng.standard_normal(g.shape)
    return x.astype(np.float32), g.astype(np.float32)


def _oracle(fn, x, y):
    xt = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    yt = torch.from_numpy(y.astype(np.float64))
    v = fn(xt, yt)
    v.backward()
    return float(v), xt.grad.numpy()


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def run_gmsd(lib, x, y, t=0.0026, alpha=0.0, lossbar=1.0):
    M, N, C, B = x.shape
    xb, yb = E.f32(x), E.f32(y)
    ws = E.aligned_bytes(lib.gmsd_workspace_bytes(M, N, C, B))
    out = np.zeros(1, np.float32)
    lib.gmsd_forward(M, N, C, B, 0, E.ptr(xb), E.ptr(yb), t, alpha, E.ptr(out), E.ptr(ws))
    lb = np.array([lossbar], np.float32)
    g = np.asfortranarray(np.full(x.shape, np.nan, np.float32))
    lib.gmsd_backward(M, N, C, B, 0, E.ptr(xb), E.ptr(yb), t, alpha, E.ptr(lb), E.ptr(ws), E.ptr(g))
    return float(out[0]), np.array(g)


def run_ssim(lib, x, y, taps=None, peakval=1.0, as_loss=True, outbar=1.0):
    M, N, C, B = x.shape
    xb, yb = E.f32(x), E.f32(y)
    ws = E.aligned_bytes(lib.ssim_workspace_bytes(M, N, C, B, taps, True))
    out = np.zeros(1, np.float32)
    lib.ssim_forward(M, N, C, B, 0, E.ptr(xb), E.ptr(yb), taps, peakval, as_loss, E.ptr(out), E.ptr(ws), True)
    ob = np.array([outbar], np.float32)
    g = np.asfortranarray(np.full(x.shape, np.nan, np.float32))
    lib.ssim_backward(M, N, C, B, 0, E.ptr(xb), E.ptr(yb), taps, as_loss, E.ptr(ob), E.ptr(ws), E.ptr(g))
    return float(out[0]), np.array(g)


# sizes straddle the 64x32 tile: smaller than a tile, ragged, multi-tile
@pytest.mark.parametrize("M,N,C,B", [(16, 12, 3, 2), (70, 37, 1, 2), (128, 64, 3, 1), (5, 3, 2, 1)])
def test_gmsd_emu_matches_oracle(emu, M, N, C, B):
    x, y = _images(M, N, C, B, 7 + M)
    v, g = run_gmsd(emu, x, y)
    vo, go = _oracle(LO.gmsd, x, y)
    assert abs(v - vo) <= 1e-5 * abs(vo)          # tolerance: rel 1e-5 (fp32 maps, fp64 sums)
    assert _rel(g, go) <= 2e-5


def test_gmsd_emu_alpha_and_cotangent(emu):
    x, y = _images(40, 33, 3, 2, 3)
    v, g = run_gmsd(emu, x, y, t=0.01, alpha=0.5, lossbar=-2.5)
    vo, go = _oracle(lambda a, b: LO.gmsd(a, b, 0.01, 0.5), x, y)
    assert abs(v - vo) <= 1e-5 * abs(vo)
    assert _rel(g, -2.5 * go) <= 2e-5


@pytest.mark.parametrize("M,N,C,B", [(24, 20, 3, 2), (45, 70, 1, 2), (64, 64, 3, 1), (11, 11, 1, 1)])
def test_ssim_emu_matches_oracle(emu, M, N, C, B):
    x, y = _images(M, N, C, B, 11 + N)
    v, g = run_ssim(emu, x, y)
    vo, go = _oracle(LO.ssim_loss, x, y)
    assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
    assert _rel(g, go) <= 5e-5                     # sigma^2 = E[x^2] - mu^2 cancels in fp32


def test_ssim_emu_fast_box_and_value_mode(emu):
    x, y = _images(30, 26, 3, 2, 5)
    taps = [0.2] * 5
    v, g = run_ssim(emu, x, y, taps=taps)
    vo, go = _oracle(lambda a, b: LO.ssim_loss_fast(a, b, 5), x, y)
    assert abs(v - vo) <= 1e-5 * abs(vo)
    assert _rel(g, go) <= 5e-5
    v2, g2 = run_ssim(emu, x, y, as_loss=False, outbar=3.0, peakval=2.0)
    vo2, go2 = _oracle(lambda a, b: LO.ssim(a, b, None, 2.0), x, y)
    assert abs(v2 - vo2) <= 1e-5 * abs(vo2)
    assert _rel(g2, 3.0 * go2) <= 5e-5


def test_loss_argument_errors(emu):
    with pytest.raises(Exception):
        emu.gmsd_workspace_bytes(0, 4, 1, 1)
    with pytest.raises(Exception):
        emu.ssim_workspace_bytes(8, 8, 1, 1, None, True)       # 11-tap window larger than the image
    with pytest.raises(Exception):
        emu.ssim_workspace_bytes(64, 64, 1, 1, [1 / 13] * 13, True)   # more than 11 taps
