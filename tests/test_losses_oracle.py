"""CPU tests pinning the loss oracle (oracle/losses_oracle.py) by self-consistency: the reference holds no test
or golden value for gmsd.jl / ssim.jl (parity unpinned upstream)."""
import torch

from oracle import losses_oracle as LO

DT = torch.float64


def _pair(M=20, N=24, C=3, B=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(M, N, C, B, dtype=DT, generator=g), torch.rand(M, N, C, B, dtype=DT, generator=g)


def test_gmsd_literal_equals_roll_formulation():
    x, y = _pair()
    assert abs(float(LO.gmsd(x, y)) - float(LO.gmsd_roll(x, y))) < 1e-12
    assert abs(float(LO.gmsd(x, y, 0.01, 0.5)) - float(LO.gmsd_roll(x, y, 0.01, 0.5))) < 1e-12


def test_known_answers():
    x, y = _pair(seed=1)
    assert float(LO.gmsd(x, x.clone())) == 0.0                 # gms == 1 everywhere
    assert abs(float(LO.ssim(x, x.clone())) - 1.0) < 1e-12
    assert abs(float(LO.ssim(x, y)) - float(LO.ssim(y, x))) < 1e-12   # symmetric in its arguments
    assert abs(float(LO.gmsd(x, y)) - float(LO.gmsd(y, x))) < 1e-12   # alpha = 0: symmetric
    c = torch.full_like(x, 0.3)
    assert float(LO.gmsd(c, y)) >= 0.0


def test_gmsd_circular_shift_invariance():
    x, y = _pair(seed=2)
    a = LO.gmsd(x, y)
    b = LO.gmsd(torch.roll(x, (3, 7), (0, 1)), torch.roll(y, (3, 7), (0, 1)))
    assert abs(float(a) - float(b)) < 1e-12


def test_sobel_orientation():
    """imgrads' first output differentiates along dim 1 (rows), the second along dim 2 (iqa_utils.jl:15-20)."""
    M, N = 12, 10
    ramp1 = torch.arange(M, dtype=DT).reshape(M, 1, 1, 1).expand(M, N, 1, 1).contiguous()
    gx, gy = LO.imgrads(ramp1)
    assert torch.allclose(gx[2:-2, 2:-2], torch.full_like(gx[2:-2, 2:-2], 1.0))   # (1+2+1)/8 * (x[i+1]-x[i-1]) = 1
    assert torch.allclose(gy[2:-2, 2:-2], torch.zeros_like(gy[2:-2, 2:-2]))


def test_ssim_window_normalised_and_box():
    k = LO.ssim_kernel()
    assert abs(float(k.sum()) - 1.0) < 1e-6
    x, y = _pair(seed=3)
    assert 0.0 < float(LO.ssim_loss_fast(x, y)) < 1.0


def test_gradients_finite_difference():
    x, y = _pair(10, 12, 2, 1, seed=4)
    for fn in (LO.gmsd, LO.ssim_loss if x.shape[0] >= 11 else (lambda a, b: LO.ssim_loss_fast(a, b, 5))):
        xr = x.clone().requires_grad_(True)
        fn(xr, y).backward()
        d = torch.randn_like(x)
        eps = 1e-6
        fd = (float(fn(x + eps * d, y)) - float(fn(x - eps * d, y))) / (2 * eps)
        an = float((xr.grad * d).sum())
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an))


def test_losses_agree_with_a_scipy_formulation():
    """Independent of the oracle's NNlib restatement: SSIM through scipy.signal.convolve2d('valid') with the 11 x 11 Gaussian,
    GMSD through scipy.ndimage.correlate with wrap-around Sobel/8 stencils (written from the formulas, gmsd.jl:5-27 / ssim.jl:112-123)."""
    import numpy as np
    from scipy.ndimage import correlate
    from scipy.signal import convolve2d
    rng = np.random.default_rng(11)
    x = rng.random((24, 20, 2, 2)); y = np.clip(x + 0.1 * rng.standard_normal(x.shape), 0, 1)
    g = np.array(LO.SSIM_KERNEL); K = np.outer(g, g)
    tot = []
    for b in range(2):
        for c in range(2):
            f = lambda a: convolve2d(a, K, mode="valid")
            X, Y = x[:, :, c, b], y[:, :, c, b]
            mx, my = f(X), f(Y)
            sx, sy, sxy = f(X * X) - mx * mx, f(Y * Y) - my * my, f(X * Y) - mx * my
            tot.append(((2 * mx * my + 1e-4) * (2 * sxy + 9e-4) / ((mx * mx + my * my + 1e-4) * (sx + sy + 9e-4))).mean())
    assert abs(float(LO.ssim(torch.from_numpy(x), torch.from_numpy(y))) - np.mean(tot)) < 1e-12
    # GMSD: true convolution with SOBEL_X = [1 2 1; 0 0 0; -1 -2 -1]/8 (rows = dim 1) == correlation with the flipped stencil
    kx = np.array([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], float) / 8
    scores = []
    for b in range(2):
        maps = []
        for img in (x, y):
            m = []
            for c in range(2):
                a = img[:, :, c, b]
                gx = correlate(a, kx[::-1, ::-1], mode="wrap"); gy = correlate(a, kx.T[::-1, ::-1], mode="wrap")
                m.append(np.sqrt(gx * gx + gy * gy + 1e-16))
            maps.append(np.stack(m, -1))
        gms = (2 * maps[0] * maps[1] + 0.0026) / (maps[0] ** 2 + maps[1] ** 2 + 0.0026)
        scores.append(np.sqrt(((gms - gms.mean()) ** 2).mean()))
    assert abs(float(LO.gmsd(torch.from_numpy(x), torch.from_numpy(y))) - np.mean(scores)) < 1e-12
