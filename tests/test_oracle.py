"""CPU tests of the oracle itself (parity is unpinned upstream, so the oracle is
pinned by self-consistency, known answers and the committed golden vectors)."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from cases import make_case, psnr, rel_l2
from oracle import admm_tv_oracle as O

DT = torch.float64
HERE = os.path.dirname(os.path.abspath(__file__))


def _t(v):
    return torch.tensor([v], dtype=DT)


@pytest.mark.parametrize("kh,kw", [(3, 3), (7, 10), (4, 4), (1, 5)])
def test_H_Ht_adjoint(kh, kw):
    torch.manual_seed(0)
    M, N, P, B = 16, 20, 2, 3
    h = torch.rand(kh, kw, 1, 1, dtype=DT)
    x = torch.randn(M, N, P, B, dtype=DT)
    y = torch.randn(M, N, P, B, dtype=DT)
    Hx = O.H_forward(x, h)
    Hty = O.Ht_roll(y, h)
    assert abs(float((Hx * y).sum() - (x * Hty).sum())) < 1e-10


def test_D_Dt_adjoint_and_literal_conv():
    torch.manual_seed(1)
    x = torch.randn(12, 10, 1, 1, dtype=DT)
    t1 = torch.randn(12, 10, 1, 1, dtype=DT)
    t2 = torch.randn(12, 10, 1, 1, dtype=DT)
    d1, d2 = O.D_roll(x)
    assert abs(float((d1 * t1).sum() + (d2 * t2).sum() - (x * O.Dt_roll(t1, t2)).sum())) < 1e-10
    # channel 1 is the dim-2 difference x[i,j]-x[i,j-1] (W1 of ops.jl:52)
    assert torch.allclose(d1[:, 1:], x[:, 1:] - x[:, :-1])
    assert torch.allclose(d2[1:, :], x[1:, :] - x[:-1, :])


@pytest.mark.parametrize("iso", [False, True])
@pytest.mark.parametrize("kh,kw", [(0, 0), (5, 5), (4, 7)])
def test_literal_equals_independent_formulation(iso, kh, kw):
    y, h, _ = make_case(16, 24, 3, 2, kh, kw, 5)
    a = O.tvd_fft_cpu(y, _t(0.05), _t(0.3), h, iso, 7)
    b = O.tvd_fft_fast(y, _t(0.05), _t(0.3), h, iso, 7)
    c = O.tvd_fft_fast(y, _t(0.05), _t(0.3), h, iso, 7, hoist=False)
    assert rel_l2(a, b) < 1e-13 and rel_l2(a, c) < 1e-13


def test_known_answer_delta_psf_tiny_lambda():
    y, _, _ = make_case(16, 16, 1, 2, 0, 0, 3)
    h = torch.zeros(3, 3, 1, 1, dtype=DT)
    h[1, 1] = 1.0   # centred delta: pd = 1 -> H = identity
    x = O.tvd_fft_cpu(y, _t(1e-12), _t(1e-3), h, False, 30)
    assert rel_l2(x, y) < 1e-9


def test_known_answer_constant_image_is_fixed_point():
    y = torch.full((16, 16, 1, 1), 0.37, dtype=DT)
    h = O.gaussian_psf(5, 1.0)
    x = O.tvd_fft_cpu(y, _t(0.01), _t(0.1), h, False, 10)
    assert float((x - 0.37).abs().max()) < 1e-12


def test_known_answer_K1_closed_form():
    y, h, _ = make_case(16, 24, 1, 1, 5, 5, 9)
    rho = _t(0.2)
    x = O.tvd_fft_cpu(y, _t(0.03), rho, h, False, 1)
    _, C = O.spectral_tables(16, 24, h, rho, DT)
    ref = O.irfft12(C.reshape(9, 24, 1, 1) * O.rfft12(O.Ht_roll(y, h)), 16)
    assert rel_l2(x, ref) < 1e-13


def test_flip_convention_asymmetric_psf_restoration():
    """A flipped-kernel bug (H <-> H^T) fails this: restoring an image blurred with an asymmetric PSF
    through the reference's own H must raise the PSNR substantially."""
    M = N = 64
    g = torch.zeros(M, N, 1, 1, dtype=DT)
    g[10:30, 12:40] = 0.8
    g[35:60, 5:25] = 0.3
    g[40:50, 40:60] = 1.0
    h = torch.zeros(7, 7, 1, 1, dtype=DT)
    h[0, 0] = 0.5; h[1, 2] = 0.2; h[3, 3] = 0.2; h[6, 1] = 0.1     # strongly asymmetric
    y = O.synthetic_observation(g, h, 0.01, 11)
    x = O.tvd_fft_cpu(y, _t(0.0041), _t(0.021), h, False, 100)
    assert psnr(x, g) > psnr(y, g) + 10.0
    # with the wrong convention the restoration is visibly worse
    O.NNLIB_CONV_FLIPS_KERNEL = False
    try:
        xw = O.tvd_fft_cpu(y, _t(0.0041), _t(0.021), h, False, 100)
    finally:
        O.NNLIB_CONV_FLIPS_KERNEL = True
    assert psnr(x, g) > psnr(xw, g) + 5.0


def test_smoke_configuration_of_reference_test_script():
    """tests/admm_deconv_test.jl:19-20,76: 7x7 row PSF, λ=0.0041, ρ=0.021, aniso, 100 it."""
    y, h, g = make_case(64, 64, 3, 1, 7, 7, 21, psf="line", noise=0.0)
    x = O.tvd_fft_cpu(y, _t(0.0041), _t(0.021), h, False, 100)
    assert psnr(x, g) > psnr(y, g) + 3.0


@pytest.mark.parametrize("iso", [False, True])
def test_layer_grads_finite_difference(iso):
    y, h, _ = make_case(8, 8, 1, 2, 3, 3, 2)
    lam, rho = _t(0.03), _t(0.4)
    xbar = torch.randn(8, 8, 1, 2, dtype=DT)
    _, g = O.layer_grads(y, xbar, h, None, lam, rho, 4, iso)
    eps = 1e-7
    f = lambda l, r: float((O.tvd_fft_cpu(y, l, r, h, iso, 4) * xbar).sum())
    fd_l = (f(lam + eps, rho) - f(lam - eps, rho)) / (2 * eps)
    fd_r = (f(lam, rho + eps) - f(lam, rho - eps)) / (2 * eps)
    assert abs(fd_l - float(g["lam"])) < 1e-5 * max(1.0, abs(fd_l))
    assert abs(fd_r - float(g["rho"])) < 1e-5 * max(1.0, abs(fd_r))


def test_golden_vectors_match_oracle():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
    assert len(files) >= 6
    for f in files:
        d = np.load(f, allow_pickle=False)
        y = torch.from_numpy(d["y"]).double()
        h = torch.from_numpy(d["h"]).double() if "h" in d else None
        lam, rho = _t(float(d["lam"])), _t(float(d["rho"]))
        bias = _t(float(d["bias"])) if "bias" in d else None
        out, _ = O.admm_layer(y, h, bias, lam, rho, int(d["iters"]), bool(d["iso"]), float(d["creg"]), str(d["act"]))
        assert rel_l2(out, torch.from_numpy(d["x"])) < 1e-12, f


@pytest.mark.parametrize("iso", [False, True])
@pytest.mark.parametrize("kh,kw", [(0, 0), (5, 4)])
def test_teacher_forced_recursion_equals_autograd(iso, kh, kw):
    """The hand-derived adjoint (oracle/teacher_forced.py) fed its own fp64 states
    reproduces torch.autograd through the literal restatement to round-off."""
    from oracle import teacher_forced as TF
    y, h, _ = make_case(16, 32, 3, 2, kh, kw, 12)
    lam, rho = _t(0.03), _t(0.4)
    xbar = torch.randn(16, 32, 3, 2, dtype=DT)
    K = 6
    _, g = O.layer_grads(y, xbar, h, None, lam, rho, K, iso)
    x, vs = TF.forward_states(y, lam, rho, h, iso, K)
    t = TF.backward(xbar, y, lam, rho, h, iso, K, vs)
    assert rel_l2(t["x"], g["x"]) < 1e-12
    assert abs(float(t["lam"]) - float(g["lam"])) < 1e-9 * max(1, abs(float(g["lam"])))
    assert abs(float(t["rho"]) - float(g["rho"])) < 1e-9 * max(1, abs(float(g["rho"])))
    if h is not None:
        assert rel_l2(t["weight"], g["weight"]) < 1e-12


@pytest.mark.parametrize("iso", [False, True])
def test_per_iteration_parameters_teacher_forced_equals_autograd(iso):
    """EXTENSION: one (lambda_k, rho_k) per unrolled iteration.  The hand adjoint recursion with
    per-iteration entries equals fp64 autograd through the roll/spectral formulation; with all entries equal both
    reduce to the reference recursion."""
    import torch
    from oracle import admm_tv_oracle as O
    from oracle import teacher_forced as TF
    from cases import make_case
    K = 6
    y, h, g = make_case(24, 16, 2, 2, 3, 5, 17)
    lam = torch.tensor([0.004, 0.006, 0.003, 0.008, 0.005, 0.007], dtype=torch.float64)
    rho = torch.tensor([0.02, 0.05, 0.03, 0.04, 0.025, 0.06], dtype=torch.float64)
    lv, rv, hv, yv = (t.clone().requires_grad_(True) for t in (lam, rho, h, y))
    x = O.tvd_fft_fast(yv, lv, rv, hv, iso, K)
    xbar = torch.from_numpy(__import__("numpy").random.default_rng(0).standard_normal(tuple(x.shape)))
    gy, gl, gr, gh = torch.autograd.grad(x, [yv, lv, rv, hv], grad_outputs=xbar)
    _, st = TF.forward_states(y, lam, rho, h, iso, K)
    tf = TF.backward(xbar, y, lam, rho, h, iso, K, st)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(tf["x"], gy) < 1e-12 and rel(tf["weight"], gh) < 1e-11
    assert rel(tf["lam"][:-1], gl[:-1]) < 1e-11 and float(gl[-1]) == 0.0 and float(tf["lam"][-1]) == 0.0   # last z-update is dead
    assert rel(tf["rho"], gr) < 1e-11
    # equal entries == the shared-parameter recursion
    xs = O.tvd_fft_fast(y, lam[:1], rho[:1], h, iso, K)
    xe = O.tvd_fft_fast(y, lam[:1].repeat(K), rho[:1].repeat(K), h, iso, K)
    assert torch.equal(xs, xe)


def test_restated_primitives_agree_with_independent_libraries():
    """The oracle is unpinned upstream (no Julia here), so its building blocks are checked against implementations it does
    not share code with: NNlib.conv (a TRUE convolution, kernel flipped) vs scipy.signal.convolve2d('valid'), pad_circular vs
    numpy.pad('wrap'), NNlib.pad_symmetric vs numpy.pad('symmetric'), rfft over dims (1,2) with dim 1 halved vs numpy.fft."""
    from scipy.signal import convolve2d
    from oracle import losses_oracle as LO
    rng = np.random.default_rng(3)
    x = rng.standard_normal((9, 7, 2, 2)); w = rng.standard_normal((4, 3, 1, 2))          # (H,W,C,B), (kh,kw,1,C): groups = C
    out = O.nnlib_conv(torch.from_numpy(x), torch.from_numpy(w), groups=2).numpy()
    for c in range(2):
        for b in range(2):
            assert np.allclose(out[:, :, c, b], convolve2d(x[:, :, c, b], w[:, :, 0, c], mode="valid"), atol=1e-12)
    pads = (2, 1, 0, 3)                                                                    # (d1_lo, d1_hi, d2_lo, d2_hi)
    npad = ((pads[0], pads[1]), (pads[2], pads[3]), (0, 0), (0, 0))
    assert np.array_equal(O.pad_circular(torch.from_numpy(x), pads).numpy(), np.pad(x, npad, mode="wrap"))
    assert np.array_equal(LO.pad_symmetric(torch.from_numpy(x), pads).numpy(), np.pad(x, npad, mode="symmetric"))
    X = O.rfft12(torch.from_numpy(x)).numpy()
    ref = np.fft.fft(np.fft.rfft(x, axis=0), axis=1)                                        # halve dim 1, full transform along dim 2
    assert X.shape == (9 // 2 + 1, 7, 2, 2) and np.allclose(X, ref, atol=1e-12)
    assert np.allclose(O.irfft12(torch.from_numpy(X), 9).numpy(), x, atol=1e-12)
