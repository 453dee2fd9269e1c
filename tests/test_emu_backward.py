"""CPU tests of the backward kernel sources through the emulation (see test_emu_forward.py).
Arithmetic parity is checked teacher-forced (the fp64 adjoint replays the checkpointed states of
the emulated forward); end-to-end agreement with fp64 autograd is checked where no mask flips."""
import glob
import os

import numpy as np
import pytest
import torch

import harness
from cases import make_case, rel_l2
from parity import T, check_backward

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def be(emu):
    return harness.EmuBackend(emu)


@pytest.mark.parametrize(
    "M,N,P,B,kh,kw,K,act,bias,flags",
    [
        (32, 32, 1, 2, 0, 0, 1, "identity", None, 0),
        (32, 32, 1, 2, 0, 0, 3, "identity", None, 0),
        (32, 64, 3, 1, 5, 4, 4, "identity", None, 0),
        (64, 32, 1, 2, 3, 3, 5, "relu1", 0.02, 0),
        (32, 32, 2, 1, 3, 3, 3, "relu", None, 2),       # train.jl:10 @nograd repeat variant
    ],
)
def test_emu_backward_aniso(be, M, N, P, B, kh, kw, K, act, bias, flags):
    y, h, _ = make_case(M, N, P, B, kh, kw, 7 + M + K)
    xbar = torch.from_numpy(np.random.default_rng(K).standard_normal((M, N, P, B)))
    check_backward(be, y, h, 0.05, 0.3, False, K, xbar, act, bias, 0.0, flags, tol=1e-5, tol_scalar=1e-4, tol_e2e=1e-3)


def test_emu_clamp_masks_gate_gradients(be):
    """deconv_admm.jl:216-219 under Zygote: clamp passes the gradient only inside the range."""
    y, h, _ = make_case(32, 32, 1, 2, 3, 3, 4)
    hh = h.clone()
    hh[0, 0] = -0.5
    hh[1, 1] = 1.7
    xbar = torch.from_numpy(np.random.default_rng(0).standard_normal((32, 32, 1, 2)))
    f = be.forward(y.numpy(), 0.01, 0.3, hh.numpy()[:, :, 0, 0], False, 3, creg=0.05, want_ckpt=True)   # lambda < creg
    g = be.backward(f, xbar.numpy())
    assert f["lam"].get()[0] == np.float32(0.05)
    assert g["lambar"][0] == 0.0 and g["rhobar"][0] != 0.0
    assert g["hbar"][0, 0] == 0.0 and g["hbar"][1, 1] == 0.0 and g["hbar"][2, 2] != 0.0
    check_backward(be, y, hh, 0.2, 0.3, False, 3, xbar, creg=0.05)


def test_emu_golden_backward(be):
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "aniso_*.npz"))):
        d = np.load(f)
        if d["y"].shape[0] > 64:
            continue
        y = torch.from_numpy(d["y"]).double()
        h = torch.from_numpy(d["h"]).double() if "h" in d else None
        r = check_backward(be, y, h, float(d["lam"]), float(d["rho"]), False, int(d["iters"]), torch.from_numpy(d["xbar"]),
                           str(d["act"]), float(d["bias"]) if "bias" in d else None, float(d["creg"]), tol=1e-5, tol_scalar=1e-4)
        print(os.path.basename(f), r)


@pytest.mark.parametrize("iso_flag", [16, 32])   # ADMMTV_FLAG_ISO_PRECOMPUTE / ADMMTV_FLAG_ISO_INLINE: both code paths
@pytest.mark.parametrize("M,N,P,B,kh,kw,K", [(32, 32, 1, 4, 0, 0, 5), (32, 64, 3, 1, 5, 4, 4)])
def test_emu_backward_iso(be, M, N, P, B, kh, kw, K, iso_flag):
    y, h, _ = make_case(M, N, P, B, kh, kw, 50 + M + K)
    xbar = torch.from_numpy(np.random.default_rng(K).standard_normal((M, N, P, B)))
    check_backward(be, y, h, 0.05, 0.3, True, K, xbar, flags=1 | iso_flag, tol=1e-5, tol_scalar=1e-4, tol_e2e=1e-4)


@pytest.mark.parametrize("iso", [False, True])
def test_emu_backward_mixed_radix(be, iso):
    y, h, _ = make_case(96, 160, 3, 1, 5, 4, 7)
    xbar = torch.from_numpy(np.random.default_rng(1).standard_normal((96, 160, 3, 1)))
    check_backward(be, y, h, 0.05, 0.3, iso, 3, xbar, tol=1e-5, tol_scalar=2e-4)


# ---- sizes without a register-FFT plan (generic_kernels.cuh) ---------------------------------------------------------
@pytest.mark.parametrize(
    "M,N,P,B,kh,kw,K,iso,act,bias,flags",
    [(20, 24, 1, 2, 3, 3, 4, False, "identity", None, 0), (33, 17, 3, 1, 5, 4, 3, False, "relu1", 0.02, 0),
     (7, 5, 1, 1, 0, 0, 3, False, "identity", None, 0), (20, 24, 2, 2, 3, 3, 4, True, "identity", None, 1 | 16),
     (31, 18, 1, 3, 0, 0, 3, True, "identity", None, 1 | 32), (30, 32, 1, 1, 3, 3, 3, False, "relu", None, 2),
     (48, 32, 1, 2, 3, 3, 3, False, "identity", None, 0), (48, 32, 2, 1, 3, 3, 3, True, "identity", None, 1 | 16),   # mixed
     (32, 224, 1, 1, 3, 3, 3, False, "identity", None, 0)],                                                          # mixed, other way
)
def test_emu_backward_generic_sizes(be, M, N, P, B, kh, kw, K, iso, act, bias, flags):
    y, h, _ = make_case(M, N, P, B, kh, kw, 9 + M)
    xbar = torch.from_numpy(np.random.default_rng(K).standard_normal((M, N, P, B)))
    check_backward(be, y, h, 0.05, 0.3, iso, K, xbar, act, bias, 0.0, flags, tol=1e-5, tol_scalar=2e-4,
                   tol_e2e=1e-4 if iso else 1e-3)


@pytest.mark.parametrize("M,N,P,B,kh,kw,K,iso", [(32, 32, 1, 2, 3, 3, 5, False), (64, 32, 3, 1, 5, 4, 4, True), (33, 17, 1, 2, 3, 3, 4, False)])
def test_emu_per_iteration_parameters(emu, M, N, P, B, kh, kw, K, iso):
    """ADMMTV_FLAG_PER_ITER_PARAMS (EXTENSION): forward vs the fp64 oracle, backward teacher-forced with one
    lambdabar / rhobar per iteration; equal entries reproduce the shared-parameter call bit for bit."""
    import numpy as np
    import harness
    from parity import check_backward, check_forward
    be = harness.EmuBackend(emu)
    y, h, g = make_case(M, N, P, B, kh, kw, 1300 + M)
    rng = np.random.default_rng(5)
    lam = (0.004 * (1 + rng.random(K))).astype(np.float32); rho = (0.02 * (1 + 2 * rng.random(K))).astype(np.float32)
    check_forward(be, y, h, lam, rho, iso, K)
    xbar = 2.0 * (y - g) / y.numel() * 1e3
    r = check_backward(be, y, h, lam, rho, iso, K, xbar, flags=1, tol=1e-5, tol_scalar=2e-4)
    print(r)
    a = be.forward(y.numpy(), float(lam[0]), float(rho[0]), h.numpy()[:, :, 0, 0], iso, K, flags=1)
    b = be.forward(y.numpy(), np.full(K, lam[0]), np.full(K, rho[0]), h.numpy()[:, :, 0, 0], iso, K, flags=1)
    assert np.array_equal(a["x"].get(), b["x"].get())


@pytest.mark.parametrize("M,N,P,B,act", [(32, 64, 3, 1, "relu1"), (33, 17, 1, 2, "identity")])
def test_emu_backward_mse_equals_backward_with_explicit_cotangent(emu, M, N, P, B, act):
    """admmtv_backward_mse (MSE pullback seed formed inside the first kernel) == admmtv_backward fed xbar = 2 (x - t) / numel."""
    be = harness.EmuBackend(emu)
    y, h, g = make_case(M, N, P, B, 3, 3, 700 + M)
    f = be.forward(y.numpy(), 0.0041, 0.021, h.numpy()[:, :, 0, 0], False, 4, act=act, bias=0.02, want_ckpt=True)
    x = f["x"].get().astype(np.float64)
    t = g.numpy()
    a = be.backward(f, 2.0 * (x - t.astype(np.float32)) / x.size)
    b = be.backward_mse(f, t)
    assert abs(b["loss"] - float(((x - t.astype(np.float32)) ** 2).mean())) <= 1e-6 * b["loss"]
    for k in ("ybar", "hbar", "lambar", "rhobar", "biasbar"):
        assert float(np.abs(a[k] - b[k]).max()) <= 2e-6 * float(np.abs(a[k]).max()) + 1e-12, k
