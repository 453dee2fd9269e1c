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
    # reduction = sum over the per-image scores (gmsd.jl:13 takes any reduction; mean and sum are fused)
    v2, g2 = _gpu(lambda a, b: A.gmsd(a, b, 0.01, 0.5, torch.sum), x, y)
    assert abs(v2 - 2 * vo) <= 1e-5 * abs(2 * vo) and _rel(g2, 2 * go) <= 2e-5


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


def _gpu_both(fn, x, y):
    d = torch.device("cuda:0")
    xt = A.from_julia(torch.from_numpy(x)).to(d).requires_grad_(True)
    yt = A.from_julia(torch.from_numpy(y)).to(d).requires_grad_(True)
    v = fn(xt, yt)
    v.backward()
    torch.cuda.synchronize()
    return float(v.detach()), A.to_julia(xt.grad.cpu()).numpy(), A.to_julia(yt.grad.cpu()).numpy()


@pytest.mark.parametrize("M,N,C,B,L1,L2", [(24, 20, 3, 2, 5, 5), (100, 70, 1, 2, 7, 3), (128, 96, 3, 1, 11, 11), (64, 40, 2, 1, 1, 4)])
def test_ssim_arbitrary_window(M, N, C, B, L1, L2):
    """ssim.jl:84 `kernel_ref`: any 2-D window (here full rank, non-symmetric) through the rank-R separable kernels."""
    from test_emu_losses import _window
    x, y = _images(M, N, C, B, 41 + L1)
    W = _window(L1, L2, L1 * 16 + L2)                      # W[a, b], the Julia (L1, L2) array
    kt = torch.from_numpy(W.T.copy())                       # torch view (L2, L1)
    v, g = _gpu(lambda a, b: A.ssim_loss(a, b, kt), x, y)
    vo, go = _oracle(lambda a, b: LO.ssim_loss(a, b, torch.from_numpy(W).reshape(L1, L2, 1, 1)), x, y)
    assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
    assert _rel(g, go) <= 5e-5


def test_ssim_per_channel_windows_and_gaussian_as_2d():
    from test_emu_losses import _window
    x, y = _images(48, 40, 3, 2, 17)
    Ws = [_window(5, 5, 100 + c) for c in range(3)]
    k4 = torch.from_numpy(np.stack([w.T for w in Ws])[:, None])          # (C, 1, L2, L1)
    v, g = _gpu(lambda a, b: A.ssim_loss(a, b, k4), x, y)
    ko = torch.from_numpy(np.stack(Ws, axis=-1)).reshape(5, 5, 1, 3)      # Julia (L1, L2, 1, C)
    vo, go = _oracle(lambda a, b: LO.ssim_loss(a, b, ko), x, y)
    assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
    assert _rel(g, go) <= 5e-5
    # the default Gaussian handed over as its 2-D outer product takes the unrolled tap kernels: same bits
    g11 = np.array(LO.SSIM_KERNEL)
    d = torch.device("cuda:0")
    xt = A.from_julia(torch.from_numpy(x)).to(d); yt = A.from_julia(torch.from_numpy(y)).to(d)
    assert float(A.ssim(xt, yt)) == float(A.ssim(xt, yt, np.outer(g11, g11)))


@pytest.mark.parametrize("kernel_length", [None, 4, 5])
def test_ssim_crop_false(kernel_length):
    """ssim.jl:104-110: same-size maps via pad_symmetric (cld / fld split for even windows) and the padding's pullback."""
    x, y = _images(70, 45, 3, 2, 23)
    if kernel_length is None:
        fn, fo = (lambda a, b: A.ssim_loss(a, b, crop=False)), (lambda a, b: LO.ssim_loss(a, b, None, 1.0, False))
    else:
        fn = lambda a, b: A.ssim_loss_fast(a, b, kernel_length, crop=False)
        fo = lambda a, b: LO.ssim_loss_fast(a, b, kernel_length, 1.0, False)
    v, g = _gpu(fn, x, y)
    vo, go = _oracle(fo, x, y)
    assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
    assert _rel(g, go) <= 5e-5


def test_pad_symmetric_bit_exact():
    from admm_deconv_b200.losses import pad_symmetric
    x = torch.randn(2, 3, 9, 7, device="cuda:0")            # (B, C, N, M)
    for pads in ((5, 5, 5, 5), (2, 1, 0, 3), (7, 7, 9, 9)):
        out = pad_symmetric(x, pads)
        ref = A.from_julia(LO.pad_symmetric(A.to_julia(x.cpu()), pads))
        assert torch.equal(out.cpu(), ref)


def test_gradients_wrt_the_target():
    """both losses are symmetric in their two images; the pullback w.r.t. the second argument runs the same kernels swapped"""
    from test_emu_losses import _oracle_both
    x, y = _images(96, 80, 3, 2, 29)
    for fn, fo, tol in ((A.ssim_loss, LO.ssim_loss, 5e-5), (A.gmsd_loss, LO.gmsd, 2e-5),
                        (lambda a, b: A.ssim_loss(a, b, crop=False), lambda a, b: LO.ssim_loss(a, b, None, 1.0, False), 5e-5)):
        v, gx, gy = _gpu_both(fn, x, y)
        vo, gxo, gyo = _oracle_both(fo, x, y)
        assert abs(v - vo) <= 1e-5 * max(abs(vo), 1e-3)
        assert _rel(gx, gxo) <= tol and _rel(gy, gyo) <= tol
    # target only
    d = torch.device("cuda:0")
    xt = A.from_julia(torch.from_numpy(x)).to(d)
    yt = A.from_julia(torch.from_numpy(y)).to(d).requires_grad_(True)
    A.ssim_loss(xt, yt).backward()
    vo, gxo, gyo = _oracle_both(LO.ssim_loss, x, y)
    assert _rel(A.to_julia(yt.grad.cpu()).numpy(), gyo) <= 5e-5
