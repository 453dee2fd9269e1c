"""GPU parity tests (run with -m gpu on the B200 box): the nvcc-built library, called through the
public API (admm_deconv_b200.ops -> ctypes -> C ABI), against the oracle and the golden fixtures.
Tolerance: relative L2 <= 1e-5 on the restored image (north_star)."""
import glob
import os

import numpy as np
import pytest
import torch

import admm_deconv_b200 as A
from cases import make_case, psnr, rel_l2
from oracle import admm_tv_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-5


def dev():
    return torch.device("cuda:0")


def run_gpu(y, h, lam, rho, iso, K):
    """y (M,N,P,B) fp64/fp32 torch -> x (M,N,P,B) fp32 cpu."""
    d = dev()
    yt = A.from_julia(y.float()).to(d)
    ht = None if h is None else A.from_julia(h.float()).to(d)
    x = A.tvd_fft(yt, torch.tensor([lam], dtype=torch.float32, device=d), torch.tensor([rho], dtype=torch.float32, device=d),
                  ht, iso, K)
    torch.cuda.synchronize()
    return A.to_julia(x.cpu())


def oracle(y, h, lam, rho, iso, K, fast=False):
    y = y.float().double()
    h = None if h is None else h.float().double()
    l = torch.tensor([lam], dtype=torch.float32).double()
    r = torch.tensor([rho], dtype=torch.float32).double()
    f = O.tvd_fft_fast if fast else O.tvd_fft_cpu
    return f(y, l, r, h, iso, K)


def test_library_is_the_native_one():
    lib = A.load()
    assert lib.path.endswith("libadmmtv.so") and os.path.exists(lib.path)
    maps = open("/proc/self/maps").read()
    assert "libadmmtv.so" in maps


@pytest.mark.parametrize(
    "M,N,P,B,kh,kw,K",
    [
        (32, 32, 1, 2, 0, 0, 1), (32, 32, 1, 1, 0, 0, 4), (32, 64, 3, 1, 5, 4, 3), (64, 32, 1, 3, 3, 3, 5),
        (128, 128, 3, 2, 9, 9, 25), (256, 256, 1, 1, 9, 9, 50), (128, 512, 1, 2, 7, 10, 10), (512, 128, 3, 1, 15, 15, 10),
    ],
)
def test_forward_aniso_vs_oracle(M, N, P, B, kh, kw, K):
    y, h, _ = make_case(M, N, P, B, kh, kw, 100 + M + N + K, psf="random")
    x = run_gpu(y, h, 0.0041, 0.021, False, K)
    xo = oracle(y, h, 0.0041, 0.021, False, K, fast=M * N > 128 * 128)
    assert rel_l2(x, xo) < TOL


@pytest.mark.parametrize("L", [32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_every_fft_length_both_dims(L):
    for (M, N) in ((L, 64), (64, L)):
        y, h, _ = make_case(M, N, 2, 1, 5, 5, L)
        x = run_gpu(y, h, 0.02, 0.1, False, 6)
        xo = oracle(y, h, 0.02, 0.1, False, 6, fast=True)
        assert rel_l2(x, xo) < TOL, (M, N)


def test_golden_forward_layer():
    """Committed fixtures (tests/golden/make_golden.py): full layer call incl. bias and activation."""
    d0 = dev()
    n = 0
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "aniso_*.npz"))):
        d = np.load(f)
        y = A.from_julia(torch.from_numpy(d["y"])).to(d0)
        h = A.from_julia(torch.from_numpy(d["h"])).to(d0) if "h" in d else None
        lam = torch.tensor([float(d["lam"])], dtype=torch.float32, device=d0)
        rho = torch.tensor([float(d["rho"])], dtype=torch.float32, device=d0)
        bias = torch.tensor([float(d["bias"])], dtype=torch.float32, device=d0) if "bias" in d else None
        x = A.admm_layer_call(y, lam, rho, h, bias, int(d["iters"]), bool(d["iso"]), str(d["act"]), float(d["creg"]))
        torch.cuda.synchronize()
        assert rel_l2(A.to_julia(x.cpu()), torch.from_numpy(d["x"])) < TOL, f
        n += 1
    assert n >= 4


def test_reference_smoke_configuration():
    """tests/admm_deconv_test.jl:19-20,67,76: (256,256,3,3), 7x7 row PSF, λ=0.0041, ρ=0.021, aniso, 100 it."""
    y, h, g = make_case(256, 256, 3, 3, 7, 7, 21, psf="line", noise=0.0)
    x = run_gpu(y, h, 0.0041, 0.021, False, 100)
    xo = oracle(y, h, 0.0041, 0.021, False, 100, fast=True)
    assert rel_l2(x, xo) < TOL
    assert psnr(x, g) > psnr(y, g) + 3.0


def test_cfg1_single_256_gaussian_50it():
    """BASELINE.json configs[0]: single 256x256 grayscale, Gaussian PSF sigma=2, 50 iterations."""
    y, h, g = make_case(256, 256, 1, 1, 9, 9, 1001, psf="gauss", noise=0.01)
    x = run_gpu(y, h, 0.0041, 0.021, False, 50)
    xo = oracle(y, h, 0.0041, 0.021, False, 50, fast=True)
    assert rel_l2(x, xo) < TOL


def test_linearity_in_y_at_full_size_when_tau_zero():
    """Size-independent property at BASELINE cfg2 plane size: with λ = 0 the solver is linear in y."""
    d0 = dev()
    torch.manual_seed(0)
    B, P, N, M = 4, 3, 512, 512
    y1 = torch.rand(B, P, N, M, device=d0)
    y2 = torch.rand(B, P, N, M, device=d0)
    h = A.from_julia(O.motion_psf(15, 0.7, 11.0).float()).to(d0)
    lam = torch.zeros(1, device=d0)
    rho = torch.tensor([0.021], device=d0)
    f = lambda y: A.tvd_fft(y, lam.clone(), rho.clone(), h.clone(), False, 10)
    a, b, c = f(y1), f(y2), f(2 * y1 - 3 * y2)
    assert rel_l2(c.cpu(), (2 * a - 3 * b).cpu()) < 1e-5


def test_planes_are_independent_and_batch_order_equivariant():
    d0 = dev()
    torch.manual_seed(1)
    y = torch.rand(5, 3, 128, 128, device=d0)
    h = A.from_julia(O.gaussian_psf(9, 2.0).float()).to(d0)
    lam = torch.tensor([0.0041], device=d0); rho = torch.tensor([0.021], device=d0)
    full = A.tvd_fft(y, lam, rho, h, False, 8)
    perm = torch.tensor([3, 0, 4, 1, 2], device=d0)
    shuf = A.tvd_fft(y[perm].contiguous(), lam, rho, h, False, 8)
    # pairing of planes changes (odd S = 15), results must not
    assert rel_l2(shuf.cpu(), full[perm].cpu()) < 1e-5
    one = A.tvd_fft(y[2:3].contiguous(), lam, rho, h, False, 8)
    assert rel_l2(one.cpu(), full[2:3].cpu()) < 1e-5


def test_host_buffer_entry_point():
    y, h, _ = make_case(64, 64, 1, 2, 5, 5, 8)
    xo = oracle(y, h, 0.02, 0.1, False, 6)
    x = A.tvd_fft_host(A.from_julia(y.float()).numpy(), 0.02, 0.1, A.from_julia(h.float()).numpy()[0, 0], False, 6)
    assert rel_l2(A.to_julia(torch.from_numpy(x)), xo) < TOL


def test_error_codes_on_bad_shapes():
    d0 = dev()
    y = torch.zeros(1, 1, 48, 64, device=d0)
    with pytest.raises(A.AdmmTvError) as e:
        A.tvd_fft(y, torch.ones(1, device=d0), torch.ones(1, device=d0), None, False, 2)
    assert e.value.code == -3


@pytest.mark.parametrize(
    "M,N,P,B,kh,kw,K",
    [(32, 32, 1, 2, 0, 0, 2), (64, 64, 3, 2, 5, 5, 15), (128, 256, 3, 2, 9, 9, 20), (256, 256, 3, 2, 15, 15, 50), (512, 512, 1, 2, 7, 7, 10)],
)
def test_forward_iso_vs_oracle(M, N, P, B, kh, kw, K):
    """Isotropic TV (BT): the norm couples every plane of the call."""
    y, h, _ = make_case(M, N, P, B, kh, kw, 500 + M + K)
    x = run_gpu(y, h, 0.0041, 0.021, True, K)
    xo = oracle(y, h, 0.0041, 0.021, True, K, fast=True)
    assert rel_l2(x, xo) < TOL


def test_golden_forward_iso():
    d0 = dev()
    n = 0
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "iso_*.npz"))):
        d = np.load(f)
        y = A.from_julia(torch.from_numpy(d["y"])).to(d0)
        h = A.from_julia(torch.from_numpy(d["h"])).to(d0) if "h" in d else None
        lam = torch.tensor([float(d["lam"])], dtype=torch.float32, device=d0)
        rho = torch.tensor([float(d["rho"])], dtype=torch.float32, device=d0)
        x = A.admm_layer_call(y, lam, rho, h, None, int(d["iters"]), True, str(d["act"]), float(d["creg"]))
        torch.cuda.synchronize()
        assert rel_l2(A.to_julia(x.cpu()), torch.from_numpy(d["x"])) < TOL, f
        n += 1
    assert n >= 2


def test_iso_couples_the_batch():
    """Reference semantics (a-9 iv): with BT an image's result depends on the other images of the call."""
    d0 = dev()
    torch.manual_seed(2)
    y = torch.rand(2, 1, 64, 64, device=d0)
    lam = torch.tensor([0.05], device=d0); rho = torch.tensor([0.3], device=d0)
    both = A.tvd_fft(y, lam, rho, None, True, 10)
    alone = A.tvd_fft(y[:1].contiguous(), lam, rho, None, True, 10)
    assert rel_l2(both[:1].cpu(), alone.cpu()) > 1e-3
