"""GPU tests of the host-buffer entry points (include/admmtv_host.h, admm_deconv_b200/host.py): the reference's
tvd_fft on a CPU Array (ops.jl:183-187) and one train.jl:49-54 step, through the two-slot pipelined session.
Checked against the device-pointer path (same kernels => bit-identical images) and the fp64 oracle."""
import numpy as np
import pytest
import torch

import admm_deconv_b200 as A
from admm_deconv_b200 import host
from cases import make_case, rel_l2
from oracle import admm_tv_oracle as O

pytestmark = pytest.mark.gpu


def _case(M, N, P, B, k, seed):
    y, h, g = make_case(M, N, P, B, k, k, seed)
    return A.from_julia(y.float()), A.from_julia(h.float()), A.from_julia(g.float()), y, h, g


@pytest.mark.parametrize("pinned", [True, False])
def test_host_forward_equals_device_path_and_oracle(pinned):
    d0 = torch.device("cuda:0")
    yt, ht, _, y, h, _ = _case(64, 128, 3, 2, 7, 11)
    lam = torch.tensor([0.0041]); rho = torch.tensor([0.021])
    s = host.HostSession(64, 128, 3, 2, 7, 7, iters=12)
    yh = yt.pin_memory() if pinned else yt
    x = s.forward(yh, lam, rho, ht.clone())
    xd = A.tvd_fft(yt.to(d0), lam.to(d0), rho.to(d0), ht.to(d0), False, 12)
    assert torch.equal(x, xd.cpu())
    xo = O.tvd_fft_fast(y.float().double(), lam.double(), rho.double(), h.float().double(), False, 12)
    assert rel_l2(A.to_julia(x), xo) < 1e-5
    s.close()


def test_host_forward_persists_the_clamp():
    """deconv_admm.jl:216-219: lambda, rho clamped to [creg, inf), weight to [0, 1], written back to the caller's arrays."""
    yt, ht, *_ = _case(32, 32, 1, 2, 3, 5)
    hh = ht.clone(); hh.view(-1)[0] = -0.5; hh.view(-1)[1] = 1.5
    lam = torch.tensor([1e-4]); rho = torch.tensor([0.5])
    s = host.HostSession(32, 32, 1, 2, 3, 3, iters=3, creg=1e-2)
    s.forward(yt, lam, rho, hh)
    assert float(lam) == pytest.approx(1e-2) and float(rho) == 0.5
    assert float(hh.view(-1)[0]) == 0.0 and float(hh.view(-1)[1]) == 1.0
    s.close()


def test_host_train_step_equals_autograd_path():
    d0 = torch.device("cuda:0")
    M, N, P, B, k, K = 64, 64, 3, 4, 5, 8
    yt, ht, gt, *_ = _case(M, N, P, B, k, 21)
    lam = torch.tensor([0.0041]); rho = torch.tensor([0.021]); bias = torch.tensor([0.01])
    s = host.HostSession(M, N, P, B, k, k, iters=K, activation="relu1", has_bias=True, training=True)
    ybar = torch.empty_like(yt)
    grads, loss = s.train_step(yt.pin_memory(), gt.pin_memory(), lam.clone(), rho.clone(), ht.clone(), bias, ybar=ybar)
    assert grads.numel() == k * k + 3 == s.ngrad
    # the same step through the device-pointer autograd path
    yd = yt.to(d0).requires_grad_(True)
    l = lam.to(d0).requires_grad_(True); r = rho.to(d0).requires_grad_(True)
    hd = ht.to(d0).requires_grad_(True); bd = bias.to(d0).requires_grad_(True)
    x = A.admm_layer_call(yd, l, r, hd, bd, K, False, "relu1", 0.0, False)
    ls = ((x - gt.to(d0)) ** 2).mean()
    ls.backward()
    ref = torch.cat([hd.grad.reshape(-1), l.grad, r.grad, bd.grad]).cpu()
    assert abs(float(loss) - float(ls)) <= 1e-6 * abs(float(ls))
    assert rel_l2(grads[: k * k], ref[: k * k]) < 1e-5
    for i in range(k * k, k * k + 3):
        assert abs(float(grads[i]) - float(ref[i])) <= 1e-5 * max(abs(float(ref[i])), 1e-6), (i, float(grads[i]), float(ref[i]))
    assert rel_l2(ybar, yd.grad.cpu()) < 1e-6
    s.close()


def test_host_pipelined_slots_match_sequential():
    """Enqueue step i+1 (other slot) before waiting for step i: results equal the one-at-a-time calls."""
    M, N, P, B, k, K = 128, 64, 1, 4, 5, 6
    lam = torch.tensor([0.0041]); rho = torch.tensor([0.021])
    batches = [_case(M, N, P, B, k, 100 + i) for i in range(5)]
    s = host.HostSession(M, N, P, B, k, k, iters=K, training=True)
    seq = [s.train_step(b[0].pin_memory(), b[2].pin_memory(), lam.clone(), rho.clone(), b[1].clone()) for b in batches]
    seq = [(g.clone(), float(l)) for g, l in seq]
    ys = [b[0].pin_memory() for b in batches]; gs = [b[2].pin_memory() for b in batches]; hs = [b[1].clone() for b in batches]
    outs = [None] * 5
    pend = {}
    for i in range(5):
        sl = i & 1
        if sl in pend:
            s.wait(sl)
        outs[i] = s.train_step_enqueue(sl, ys[i], gs[i], lam.clone(), rho.clone(), hs[i],
                                       grads=torch.empty(s.ngrad).pin_memory(), loss=torch.empty(1).pin_memory())
        pend[sl] = i
    s.wait(0); s.wait(1)
    for (g, l), (gr, lr) in zip(outs, seq):
        assert float(l) == lr
        assert rel_l2(g[: k * k], gr[: k * k]) < 1e-6     # spectral accumulators use float atomics: equal to rounding
    s.close()


def test_host_session_rejects_device_tensors():
    s = host.HostSession(32, 32, 1, 1, 0, 0, iters=2)
    with pytest.raises(RuntimeError, match="CPU tensors"):
        s.forward(torch.zeros(1, 1, 32, 32, device="cuda:0"), torch.tensor([0.1]), torch.tensor([0.1]))
    s.close()


@pytest.mark.parametrize("layout", ["BCNM", "BNMC"])
def test_host_train_step_from_8bit_samples(layout):
    """admmtv_host_train_step_enqueue_n0f8: the batch and the target travel as 8-bit samples (the dataset's format,
    base_funcs.jl:29-35 converts them on the CPU) and are converted on the device; same result as the float call on value / 255,
    pipelined over both slots."""
    M, N, P, B, k, K = 64, 96, 3, 4, 5, 6
    rng = np.random.Generator(np.random.PCG64(5))
    _, ht, *_ = _case(M, N, P, B, k, 31)
    lam = torch.tensor([0.0041]); rho = torch.tensor([0.021])
    s = host.HostSession(M, N, P, B, k, k, iters=K, training=True)
    for step in range(3):
        yu = torch.from_numpy(rng.integers(0, 256, size=(B, P, N, M), dtype=np.uint8))
        tu = torch.from_numpy(rng.integers(0, 256, size=(B, P, N, M), dtype=np.uint8))
        yf = (yu.float() / 255.0).pin_memory(); tf = (tu.float() / 255.0).pin_memory()
        g0, l0 = s.train_step(yf, tf, lam.clone(), rho.clone(), ht.clone())
        g0, l0 = g0.clone(), float(l0)
        if layout == "BNMC":
            yu, tu = yu.permute(0, 2, 3, 1).contiguous(), tu.permute(0, 2, 3, 1).contiguous()
        sl = step & 1
        g1, l1 = s.train_step_enqueue_n0f8(sl, yu.pin_memory(), tu.pin_memory(), lam.clone(), rho.clone(), ht.clone(), layout=layout)
        s.wait(sl)
        assert float(l1) == l0
        assert rel_l2(g1[: k * k], g0[: k * k]) < 1e-6 and torch.allclose(g1[k * k:], g0[k * k:], rtol=1e-5, atol=0)
    s.close()


def test_host_forward_from_8bit_samples():
    """admmtv_host_forward_enqueue_n0f8: a decoded 8-bit image goes up as bytes; same restored image as the float call on value / 255."""
    M, N, P, B, k = 96, 64, 3, 2, 7
    rng = np.random.Generator(np.random.PCG64(9))
    _, ht, *_ = _case(M, N, P, B, k, 41)
    lam = torch.tensor([0.0041]); rho = torch.tensor([0.021])
    s = host.HostSession(M, N, P, B, k, k, iters=9)
    yu = torch.from_numpy(rng.integers(0, 256, size=(B, P, N, M), dtype=np.uint8))
    x0 = s.forward((yu.float() / 255.0).pin_memory(), lam.clone(), rho.clone(), ht.clone()).clone()
    for layout, src in (("BCNM", yu), ("BNMC", yu.permute(0, 2, 3, 1).contiguous())):
        x1 = s.forward_enqueue_n0f8(1, src.pin_memory(), lam.clone(), rho.clone(), ht.clone(), layout=layout)
        s.wait(1)
        assert torch.equal(x1, x0), layout
    s.close()
