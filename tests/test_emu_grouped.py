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


@pytest.mark.parametrize("iso", [False, True])
def test_emu_per_image_psf_and_noise_level(be, iso):
    """BASELINE configs[4] semantics: every image has its own PSF and (lambda, rho); oracle = B=1 calls."""
    M, N, P, B, K = 32, 32, 1, 3, 4
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


@pytest.mark.parametrize("iso", [False, True])
def test_emu_parallel_branches_shared_input_channel_concat(be, iso):
    """net_build.jl:113-128: 5 x ADMMDeconvF2((), K, rho_i, relu1) on the same input, chcat."""
    M, N, P, B, K, G = 32, 32, 3, 2, 3, 5
    y, _, _ = make_case(M, N, P, B, 0, 0, 77)
    y = y.float().double()
    rhos = [0.05, 0.1, 0.2, 0.4, 0.8]
    lams = [0.03] * G
    x = be.forward_grouped(y.numpy(), lams, rhos, None, iso, K, groups=G, shared_input=True, concat=True, act="relu1")
    ref = torch.cat([O.ACTIVATIONS["relu1"](O.tvd_fft_cpu(y, _t(lams[g]), _t(rhos[g]), None, iso, K)) for g in range(G)], dim=2)
    assert x.shape == (M, N, G * P, B)
    assert rel_l2(T(x), ref) < 1e-5
