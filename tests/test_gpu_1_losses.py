"""GPU parity tests of the loss kernels (include/admmtv_loss.h) through the public API
(admm_deconv_b200.losses -> ctypes -> C ABI -> sm_100a kernels) against the fp64 oracle restatement of
src/metrics/gmsd.jl and src/metrics/ssim.jl.  Tolerances: value rel 1e-5, gradient rel-L2 2e-5 (GMSD) /
5e-5 (SSIM, whose variances E[x^2]-mu^2 cancel in fp32)."""
import numpy as np
import pytest
import torch

import admm_deconv_b200 as A
from oracle import losses_oracle as LO
from test_emu_losses import _images, _oracle, _rel

pytestmark = pytest.mark.gpu


def _gpu(fn, x, y, cot=1.0):
    d = torch.device("cuda:0")
    xt = A.from_julia(torch.from_numpy(x)).to(d).requires_grad_(True)
    yt = A.from_julia(torch.from_numpy(y)).to(d)
    v = fn(xt, yt)
    (v * cot).backward()
    torch.cuda.synchronize()
    return float(v.detach()), A.to_julia(xt.grad.cpu()).numpy()


@pytest.mark.parametrize("M,N,C,B", [(16, 12, 3, 2), (70, 37, 1, 2), (256, 256, 3, 2), (100, 333, 3, 1)])
def test_gmsd_vs_oracle(M, N, C, B):
    x, y = _images(M, N, C, B, 21 + M)
    v, g = _gpu(A.gmsd_loss, x, y)
    vo, go = _oracle(LO.gmsd, x, y)
    assert abs(v - vo) <= 1e-5 * abs(vo)
    assert _rel(g, go) <= 2e-5


def test_gmsd_params_and_cotangent():
    x, y = _images(64, 48, 3, 2, 5)
    v, g = _gpu(lambda a, b: A.gmsd(a, b, 0.01, 0.5), x, y, cot=-2.5)
    vo, go = _oracle(lambda a, b: LO.gmsd(a, b, 0.01, 0.5), x, y)
    assert abs(v - vo) <= 1e-5 * abs(vo)
    assert _rel(g, -2.5 * go) <= 2e-5


@pytest.mark.parametrize("M,N,C,B", [(24, 20, 3, 2), (45, 70, 1, 2), (256, 256, 3, 2), (100, 333, 3, 1)])
def test_ssim_loss_vs_oracle(M, N, C, B):
    x, y = _images(M, N, C, B, 31 + N)
    v, g = _gpu(A.ssim_loss, x, y)
    vo, go = _oracle(LO.ssim_loss, x, y)
    assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
    assert _rel(g, go) <= 5e-5


def test_ssim_fast_and_value():
    x, y = _images(96, 80, 3, 2, 9)
    v, g = _gpu(A.ssim_loss_fast, x, y)
    vo, go = _oracle(lambda a, b: LO.ssim_loss_fast(a, b, 5), x, y)
    assert abs(v - vo) <= 1e-5 * abs(vo)
    assert _rel(g, go) <= 5e-5
    v2, g2 = _gpu(lambda a, b: A.ssim(a, b, None, 2.0), x, y, cot=3.0)
    vo2, go2 = _oracle(lambda a, b: LO.ssim(a, b, None, 2.0), x, y)
    assert abs(v2 - vo2) <= 1e-5 * abs(vo2)
    assert _rel(g2, 3.0 * go2) <= 5e-5


def test_known_answers():
    d = torch.device("cuda:0")
    x = torch.rand(2, 3, 64, 64, device=d)
    assert abs(float(A.ssim(x, x.clone())) - 1.0) < 1e-6
    assert float(A.gmsd(x, x.clone())) < 1e-6
    # GMSD uses circular padding: invariant under a joint circular shift
    y = torch.rand(2, 3, 64, 64, device=d)
    a = float(A.gmsd(x, y))
    b = float(A.gmsd(torch.roll(x, (5, 9), (2, 3)), torch.roll(y, (5, 9), (2, 3))))
    assert abs(a - b) <= 1e-6 * a


def test_training_step_layer_then_loss():
    """ADMM layer -> gmsd_loss -> backward: the cotangent produced by the loss kernel feeds admmtv_backward; parameter
    gradients are compared with autograd through the fp64 oracles of both (teacher-free, so the tolerance is the
    end-to-end one of tests/test_gpu_3_backward.py: mask flips allowed)."""
    from cases import make_case
    from oracle import admm_tv_oracle as O
    M, N, P, B, K = 64, 64, 3, 2, 6
    y, h, gt = make_case(M, N, P, B, 5, 5, 77, psf="gauss")
    d = torch.device("cuda:0")
    yt = A.from_julia(y.float()).to(d)
    tg = A.from_julia(gt.float()).to(d)
    layer = A.ADMMDeconv((5, 5), K, "relu1").to(d)
    with torch.no_grad():
        layer.weight.copy_(A.from_julia(h.float()).to(d))
        layer.lam.fill_(0.0041); layer.rho.fill_(0.021)
    loss = A.gmsd_loss(layer(yt), tg)
    loss.backward()
    torch.cuda.synchronize()
    # oracle
    ho = h.float().double().requires_grad_(True)
    lo = torch.tensor([0.0041], dtype=torch.float32).double().requires_grad_(True)
    ro = torch.tensor([0.021], dtype=torch.float32).double().requires_grad_(True)
    xo = torch.clamp(O.tvd_fft_cpu(y.float().double(), lo, ro, ho, False, K), 0.0, 1.0)
    lo_v = LO.gmsd(xo, gt.float().double())
    lo_v.backward()
    assert abs(float(loss) - float(lo_v)) <= 1e-4 * abs(float(lo_v))
    gh = A.to_julia(layer.weight.grad.cpu()).double()
    assert float((gh - ho.grad).norm() / ho.grad.norm()) <= 5e-3
    assert abs(float(layer.lam.grad) - float(lo.grad)) <= 5e-3 * abs(float(lo.grad))
