"""Grouped calls (EXTENSION) through the emulation: the result must equal the
reference semantics applied once per group."""
import numpy as np
import pytest
import torch

import harness
from cases import make_case, rel_l2
from oracle import admm_tv_oracle as O

T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()


@pytest.fixture(scope="module")
def be(emu):
    return harness.EmuBackend(emu)


def _t(v):
    return torch.tensor([v], dtype=torch.float32).double()


@pytest.mark.parametrize("MN", [(32, 32), (18, 21)])   # tuned kernels / generic-size kernels
@pytest.mark.parametrize("iso", [False, True])
def test_emu_per_image_psf_and_noise_level(be, iso, MN):
    """BASELINE configs[4] semantics: every image has its own PSF and (lambda, rho); oracle = B=1 calls."""
    (M, N), P, B, K = MN, 1, 3, 4
    ys, hs, lams, rhos, ref = [], [], [], [], []
    for b in range(B):
        y, h, _ = make_case(M, N, P, 1, 3, 5, 900 + b)
        ys.append(y.float().double()); hs.append(h.float().double())
        lams.append(0.02 * (b + 1)); rhos.append(0.1 * (b + 2))
        ref.append(O.tvd_fft_cpu(ys[-1], _t(lams[-1]), _t(rhos[-1]), hs[-1], iso, K))
    y = torch.cat(ys, dim=3)
    h = torch.cat([hh[:, :, :, 0] for hh in hs], dim=2)          # (kh,kw,G)
    x = be.forward_grouped(y.numpy(), lams, rhos, h.numpy(), iso, K, groups=B)
    assert rel_l2(T(x), torch.cat(ref, dim=3)) < 1e-5


@pytest.mark.parametrize("MN", [(32, 32), (18, 21)])   # tuned kernels / generic-size kernels
@pytest.mark.parametrize("iso", [False, True])
def test_emu_parallel_branches_shared_input_channel_concat(be, iso, MN):
    if iso and MN != (32, 32):
        pytest.skip("generic-size grouped isotropic is covered by the per-image forward test (keeps the CPU suite short)")
    """net_build.jl:113-128: 5 x ADMMDeconvF2((), K, rho_i, relu1) on the same input, chcat."""
    (M, N), P, B, K, G = MN, 3, 2, 3, 5
    y, _, _ = make_case(M, N, P, B, 0, 0, 77)
    y = y.float().double()
    rhos = [0.05, 0.1, 0.2, 0.4, 0.8]
    lams = [0.03] * G
    x = be.forward_grouped(y.numpy(), lams, rhos, None, iso, K, groups=G, shared_input=True, concat=True, act="relu1")
    ref = torch.cat([O.ACTIVATIONS["relu1"](O.tvd_fft_cpu(y, _t(lams[g]), _t(rhos[g]), None, iso, K)) for g in range(G)], dim=2)
    assert x.shape == (M, N, G * P, B)
    assert rel_l2(T(x), ref) < 1e-5


@pytest.mark.parametrize("MN", [(32, 32), (18, 21)])   # tuned kernels / generic-size kernels
@pytest.mark.parametrize("iso", [False, True])
def test_emu_grouped_backward_per_image(be, iso, MN):
    if iso and MN != (32, 32):
        pytest.skip("generic-size grouped isotropic is covered by the per-image forward test (keeps the CPU suite short)")
    """Grouped backward == the single-call backward run group by group (same library, same inputs)."""
    (M, N), P, B, K = MN, 1, 3, 4
    ys, hs = [], []
    lams, rhos = [0.02, 0.04, 0.06], [0.2, 0.3, 0.4]
    for b in range(B):
        y, h, _ = make_case(M, N, P, 1, 3, 5, 900 + b)
        ys.append(y.float().double()); hs.append(h.float().double())
    y = torch.cat(ys, dim=3)
    h = torch.cat([hh[:, :, :, 0] for hh in hs], dim=2)
    xbar = torch.from_numpy(np.random.default_rng(1).standard_normal((M, N, P, B)))
    f = be.forward_grouped(y.numpy(), lams, rhos, h.numpy(), iso, K, groups=B, want_ckpt=True)
    g = be.backward_grouped(f, xbar.numpy())
    for b in range(B):
        f1 = be.forward(ys[b].numpy(), lams[b], rhos[b], hs[b].numpy()[:, :, 0, 0], iso, K, flags=1, want_ckpt=True)
        g1 = be.backward(f1, xbar[..., b:b + 1].numpy())
        assert rel_l2(T(g["ybar"][..., b:b + 1]), T(g1["ybar"])) < 2e-6
        assert rel_l2(T(g["hbar"][:, :, b]), T(g1["hbar"])) < 2e-5
        assert abs(g["lambar"][b] - g1["lambar"][0]) <= 1e-4 * max(abs(g1["lambar"][0]), 1e-3)
        assert abs(g["rhobar"][b] - g1["rhobar"][0]) <= 1e-4 * max(abs(g1["rhobar"][0]), 1e-3)


@pytest.mark.parametrize("MN", [(32, 32)])
@pytest.mark.parametrize("iso", [False, True])
def test_emu_grouped_backward_shared_input_branches(be, iso, MN):
    """The 5-branch denoiser bank: ybar sums over the branches, lambar / rhobar are per branch."""
    (M, N), P, B, K, G = MN, 3, 2, 3, 3
    y, _, _ = make_case(M, N, P, B, 0, 0, 77)
    y = y.float().double()
    rhos, lams = [0.05, 0.2, 0.8], [0.03, 0.02, 0.04]
    xbar = torch.from_numpy(np.random.default_rng(2).standard_normal((M, N, G * P, B)))
    f = be.forward_grouped(y.numpy(), lams, rhos, None, iso, K, groups=G, shared_input=True, concat=True, act="relu1",
                           want_ckpt=True)
    g = be.backward_grouped(f, xbar.numpy())
    ysum = 0
    for b in range(G):
        f1 = be.forward(y.numpy(), lams[b], rhos[b], None, iso, K, act="relu1", flags=1, want_ckpt=True)
        g1 = be.backward(f1, xbar[:, :, b * P:(b + 1) * P, :].numpy())
        ysum = ysum + T(g1["ybar"])
        assert abs(g["lambar"][b] - g1["lambar"][0]) <= 1e-4 * max(abs(g1["lambar"][0]), 1e-3)
        assert abs(g["rhobar"][b] - g1["rhobar"][0]) <= 1e-4 * max(abs(g1["rhobar"][0]), 1e-3)
    assert rel_l2(T(g["ybar"]), ysum) < 2e-6
