"""GPU parity tests of the hand-written backward, ISOTROPIC TV (BT, ops.jl:6,10) -- the reference's shipped training
configuration (configs/train_cfg.json:14 use_iso = true).

The per-pixel norm n couples every plane of the call.  On the device the plane pairs' shares of |v|^2 and <q,v> are
written with plain stores and added in a fixed order by k_iso_scale / k_iso_coef (no floating-point atomics), so the
isotropic path is bit-reproducible: the tests below assert torch.equal across repeated runs.  Teacher forcing replays
the device's checkpointed states AND its checkpointed norms (admmtv_ckpt_layout out[3]), so the fp64 adjoint takes
every gate n > tau exactly as the device took it; the reported `flips` counts the decisions on which the device's
forward and the fp64 forward disagree (they only matter to the end-to-end comparison)."""
import glob
import os

import numpy as np
import pytest
import torch

import admm_deconv_b200 as A
import harness
from cases import make_case, rel_l2
from parity import T, check_backward

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def be():
    return harness.GpuBackend(A.load())


@pytest.mark.parametrize("iso_flag", [0, 16, 32])   # the two legacy flags are accepted and ignored (include/admmtv.h)
@pytest.mark.parametrize("M,N,P,B,kh,kw,K", [(32, 32, 1, 4, 0, 0, 5), (64, 64, 3, 2, 7, 7, 10), (256, 128, 3, 2, 9, 9, 8), (512, 512, 1, 2, 5, 5, 4)])
def test_backward_iso_teacher_forced(be, M, N, P, B, kh, kw, K, iso_flag):
    y, h, g = make_case(M, N, P, B, kh, kw, 600 + M + K)
    xbar = 2.0 * (y - g) / y.numel() * 1e3
    r = check_backward(be, y, h, 0.0041, 0.021, True, K, xbar, flags=1 | iso_flag, tol=1e-5, tol_scalar=2e-4, tol_e2e=1e-3)
    print(r)


def test_golden_backward_iso(be):
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "iso_*.npz"))):
        d = np.load(f)
        y = torch.from_numpy(d["y"]).double()
        h = torch.from_numpy(d["h"]).double() if "h" in d else None
        r = check_backward(be, y, h, float(d["lam"]), float(d["rho"]), True, int(d["iters"]), torch.from_numpy(d["xbar"]),
                           str(d["act"]), None, float(d["creg"]), tol=1e-5, tol_scalar=2e-4)
        print(os.path.basename(f), r)



@pytest.mark.parametrize("M,N", [(384, 192)])
def test_backward_iso_mixed_radix(be, M, N):
    y, h, g = make_case(M, N, 3, 1, 7, 7, 800 + M + N)
    xbar = 2.0 * (y - g) / y.numel() * 1e3
    r = check_backward(be, y, h, 0.0041, 0.021, True, 6, xbar, tol=1e-5, tol_scalar=2e-4)
    print(r)


@pytest.mark.parametrize("M,N,P,B,kh,kw,K", [(225, 64, 1, 2, 5, 5, 6), (127, 131, 1, 2, 5, 5, 6), (720, 256, 1, 2, 5, 5, 5), (256, 360, 1, 2, 5, 5, 5)])
def test_backward_iso_any_size(be, M, N, P, B, kh, kw, K):
    """generic-size kernels (and the mixed tuned/generic dispatch), isotropic"""
    y, h, g = make_case(M, N, P, B, kh, kw, 900 + M + N)
    xbar = 2.0 * (y - g) / y.numel() * 1e3
    r = check_backward(be, y, h, 0.0041, 0.021, True, K, xbar, flags=1, tol=1e-5, tol_scalar=5e-4)   # rhobar is a cancelling sum; the direct prime-length DFTs add sqrt(L) rounding
    print(r)


@pytest.mark.parametrize("M,N,P,B,kh,kw,K", [(64, 64, 3, 2, 7, 7, 10), (256, 128, 3, 4, 9, 9, 8), (512, 512, 3, 8, 5, 5, 4), (100, 100, 3, 2, 7, 7, 6)])
def test_iso_forward_and_backward_are_bit_reproducible(be, M, N, P, B, kh, kw, K):
    """No floating-point atomics on the isotropic path: x, every checkpointed state and norm, and ybar are bit-identical
    run to run (the scalar / PSF gradients go through fp64 atomics and are compared to rounding only)."""
    y, h, g = make_case(M, N, P, B, kh, kw, 4200 + M)
    xbar = (2.0 * (y - g) / y.numel() * 1e3).numpy()
    runs = []
    for _ in range(3):
        f = be.forward(y.numpy(), 0.0041, 0.021, h.numpy()[:, :, 0, 0], True, K, flags=1, want_ckpt=True)
        gr = be.backward(f, xbar)
        runs.append((f["x"].get(), f["ckpt"].get(), gr))
    for x, ck, gr in runs[1:]:
        assert np.array_equal(x, runs[0][0])
        assert np.array_equal(ck, runs[0][1])
        assert np.array_equal(gr["ybar"], runs[0][2]["ybar"])
        assert abs(float(gr["lambar"][0]) - float(runs[0][2]["lambar"][0])) <= 1e-6 * abs(float(runs[0][2]["lambar"][0]))
