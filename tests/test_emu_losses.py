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


def _oracle_both(fn, x, y):
    xt = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    yt = torch.from_numpy(y.astype(np.float64)).requires_grad_(True)
    v = fn(xt, yt)
    v.backward()
    return float(v.detach()), xt.grad.numpy(), yt.grad.numpy()


def _factor(W):
    """(L1, L2) window -> the rank-R separable terms the C ABI takes (what losses._window_of does with an SVD)."""
    U, S, Vt = np.linalg.svd(np.asarray(W, np.float64))
    keep = [r for r in range(len(S)) if S[r] > 1e-7 * S[0]]
    return [list(U[:, r] * S[r]) for r in keep], [list(Vt[r]) for r in keep]


def run_ssim_window(lib, x, y, W, peakval=1.0, as_loss=True, outbar=1.0):
    M, N, C, B = x.shape
    u, v = _factor(W)
    L1, L2 = np.asarray(W).shape
    xb, yb = E.f32(x), E.f32(y)
    ws = E.aligned_bytes(lib.ssim_window_workspace_bytes(M, N, C, B, L1, L2, True))
    out = np.zeros(1, np.float32)
    lib.ssim_window_forward(M, N, C, B, 0, E.ptr(xb), E.ptr(yb), u, v, peakval, as_loss, E.ptr(out), E.ptr(ws), True)
    ob = np.array([outbar], np.float32)
    g = np.asfortranarray(np.full(x.shape, np.nan, np.float32))
    lib.ssim_window_backward(M, N, C, B, 0, E.ptr(xb), E.ptr(yb), u, v, as_loss, E.ptr(ob), E.ptr(ws), E.ptr(g))
    return float(out[0]), np.array(g)


def _window(L1, L2, seed):
    """a non-separable, non-symmetric normalised window (full rank)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    W = rng.random((L1, L2)) + 0.2
    return W / W.sum()


@pytest.mark.parametrize("M,N,C,B,L1,L2", [(24, 20, 3, 2, 5, 5), (45, 70, 1, 2, 7, 3), (40, 36, 2, 1, 11, 11), (12, 9, 1, 1, 1, 4)])
def test_ssim_emu_arbitrary_window(emu, M, N, C, B, L1, L2):
    """ssim.jl:84 takes any `kernel_ref`; a full-rank window goes through R = min(L1, L2) separable terms."""
    x, y = _images(M, N, C, B, 21 + L1)
    W = _window(L1, L2, L1 * 16 + L2)
    v, g = run_ssim_window(emu, x, y, W)
    k = torch.from_numpy(W).reshape(L1, L2, 1, 1)
    vo, go = _oracle(lambda a, b: LO.ssim_loss(a, b, k), x, y)
    assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
    assert _rel(g, go) <= 5e-5


def test_ssim_emu_window_path_equals_tap_path(emu):
    """the Gaussian given as a 2-D window (rank one) reproduces the unrolled separable kernels"""
    x, y = _images(40, 33, 3, 2, 9)
    g11 = np.array(LO.SSIM_KERNEL)
    v1, g1 = run_ssim(emu, x, y)
    v2, g2 = run_ssim_window(emu, x, y, np.outer(g11, g11))
    assert abs(v1 - v2) <= 2e-6 * abs(v1)
    assert _rel(g2, g1) <= 1e-5


@pytest.mark.parametrize("M,N,P,pads", [(9, 7, 3, (5, 5, 5, 5)), (6, 11, 2, (2, 1, 0, 3)), (4, 4, 1, (4, 4, 4, 4)), (5, 3, 1, (0, 0, 0, 0))])
def test_pad_symmetric_emu_and_adjoint(emu, M, N, P, pads):
    """NNlib pad_symmetric (ssim.jl:108-109): bit-exact against the oracle's index form; the adjoint is the exact transpose."""
    rng = np.random.Generator(np.random.PCG64(M * 31 + N))
    x = rng.standard_normal((M, N, P, 1)).astype(np.float32)
    lo1, hi1, lo2, hi2 = pads
    xb = E.f32(x)
    out = np.asfortranarray(np.full((M + lo1 + hi1, N + lo2 + hi2, P, 1), np.nan, np.float32))
    emu.pad_symmetric(M, N, P, pads, 0, E.ptr(xb), E.ptr(out))
    ref = LO.pad_symmetric(torch.from_numpy(x), pads).numpy()
    assert np.array_equal(np.array(out), ref)
    gbar = E.f32(rng.standard_normal(out.shape))
    xbar = np.asfortranarray(np.full(x.shape, np.nan, np.float32))
    emu.pad_symmetric_adjoint(M, N, P, pads, 0, E.ptr(gbar), E.ptr(xbar))
    xt = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    (LO.pad_symmetric(xt, pads) * torch.from_numpy(np.array(gbar, np.float64))).sum().backward()
    assert np.allclose(np.array(xbar), xt.grad.numpy(), rtol=1e-6, atol=1e-6)


def test_ssim_emu_crop_false(emu):
    """ssim.jl:104-110: crop=false = pad_symmetric by (cld(L-1,2), fld(L-1,2)) + the valid-size call; the gradient folds back
    through the adjoint of the padding."""
    M, N, C, B = 20, 17, 2, 2
    x, y = _images(M, N, C, B, 4)
    for taps, kern in ((None, None), ([0.25] * 4, LO.ssim_kernel(torch.float64, 4))):
        L = 11 if taps is None else len(taps)
        pads = (-(-(L - 1) // 2), (L - 1) // 2, -(-(L - 1) // 2), (L - 1) // 2)
        Mp, Np = M + L - 1, N + L - 1
        xp = np.asfortranarray(np.empty((Mp, Np, C, B), np.float32)); yp = np.asfortranarray(np.empty((Mp, Np, C, B), np.float32))
        xb, yb = E.f32(x), E.f32(y)
        emu.pad_symmetric(M, N, C * B, pads, 0, E.ptr(xb), E.ptr(xp))
        emu.pad_symmetric(M, N, C * B, pads, 0, E.ptr(yb), E.ptr(yp))
        v, gp = run_ssim(emu, np.array(xp), np.array(yp), taps=taps)
        g = np.asfortranarray(np.empty(x.shape, np.float32))
        gpb = E.f32(gp)
        emu.pad_symmetric_adjoint(M, N, C * B, pads, 0, E.ptr(gpb), E.ptr(g))
        vo, go = _oracle(lambda a, b: LO.ssim_loss(a, b, kern, 1.0, False), x, y)
        assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
        assert _rel(np.array(g), go) <= 5e-5


def test_target_gradient_is_the_swapped_call(emu):
    """both losses are symmetric in (x, y): d/dy through the same kernels with the images swapped"""
    x, y = _images(30, 26, 3, 2, 12)
    _, gy = run_ssim(emu, y, x)
    vo, gxo, gyo = _oracle_both(LO.ssim_loss, x, y)
    assert _rel(gy, gyo) <= 5e-5
    _, gy = run_gmsd(emu, y, x)
    vo, gxo, gyo = _oracle_both(LO.gmsd, x, y)
    assert _rel(gy, gyo) <= 2e-5


def test_loss_argument_errors(emu):
    with pytest.raises(Exception):
        emu.pad_symmetric(4, 4, 1, (5, 0, 0, 0), 0, 1, 1)      # pad larger than the array (NNlib rejects it too)
    with pytest.raises(Exception):
        emu.ssim_window_workspace_bytes(8, 8, 1, 1, 12, 3, True)   # window larger than 11
    with pytest.raises(Exception):
        emu.gmsd_workspace_bytes(0, 4, 1, 1)
    with pytest.raises(Exception):
        emu.ssim_workspace_bytes(8, 8, 1, 1, None, True)       # 11-tap window larger than the image
    with pytest.raises(Exception):
        emu.ssim_workspace_bytes(64, 64, 1, 1, [1 / 13] * 13, True)   # more than 11 taps
