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


def test_tma_pipelined_dim2_pass_full_ring(monkeypatch=None):
    """N = 2048 runs the TMA-pipelined dim-2 kernel (kernels_tma.cuh: 2-D box copies, mbarrier ring of 3 buffers, two
    compute groups per persistent block).  2 x 2048 x 2048: 512 tiles over <= 148 blocks, so every buffer of a ring is
    re-armed and reused; plus a short plane (fewer tiles than SMs) and an odd number of plane pairs."""
    for (M, N, P, B, K) in ((2048, 2048, 1, 2, 3), (32, 2048, 1, 1, 4), (128, 2048, 3, 1, 3)):
        y, h, _ = make_case(M, N, P, B, 5, 5, 77 + M)
        x = run_gpu(y, h, 0.02, 0.1, False, K)
        xo = oracle(y, h, 0.02, 0.1, False, K, fast=True)
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
    y = torch.zeros(1, 1, 48, 4100, device=d0)      # dim 1 (M) = 4100 > 4096
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


def test_grouped_per_image_psf_and_noise_level():
    """BASELINE configs[4] semantics (EXTENSION): 128x128 images, each with its own 9x9 motion PSF
    and (lambda, rho); must equal the reference semantics run once per image (oracle, B = 1) and the same
    library called image by image."""
    import math
    d0 = dev()
    B, K = 24, 20
    rng = np.random.default_rng(5)
    ys, hs, lams, rhos = [], [], [], []
    for b in range(B):
        h = O.motion_psf(9, float(rng.uniform(0, math.pi)), float(rng.uniform(5, 9)))
        g = O.synthetic_truth(128, 128, 1, 1, 4000 + b)
        sig = [0.005, 0.01, 0.02, 0.04][b % 4]
        ys.append(O.synthetic_observation(g, h, sig, 4000 + b).float()); hs.append(h.float())
        lams.append(0.2 * sig); rhos.append(1.0 * sig)
    y = A.from_julia(torch.cat(ys, dim=3)).to(d0)                               # (B,1,128,128)
    h = torch.stack([A.from_julia(hh)[0] for hh in hs]).to(d0).contiguous()      # (B,1,9,9)
    lam = torch.tensor(lams, dtype=torch.float32, device=d0); rho = torch.tensor(rhos, dtype=torch.float32, device=d0)
    x = A.tvd_fft_grouped(y, lam, rho, h, False, K, groups=B)
    torch.cuda.synchronize()
    for b in range(B):
        one = A.tvd_fft(y[b:b + 1].contiguous(), lam[b:b + 1].clone(), rho[b:b + 1].clone(), h[b:b + 1].contiguous(), False, K)
        assert rel_l2(x[b:b + 1].cpu(), one.cpu()) < 1e-5, b
    for b in (0, 7, 23):
        xo = O.tvd_fft_fast(ys[b].double(), torch.tensor([lams[b]], dtype=torch.float32).double(),
                            torch.tensor([rhos[b]], dtype=torch.float32).double(), hs[b].double(), False, K)
        assert rel_l2(A.to_julia(x[b:b + 1].cpu()), xo) < TOL, b


@pytest.mark.parametrize("iso", [False, True])
def test_grouped_parallel_branches_of_get_denoiser(iso):
    """net_build.jl:113-128: 5 x ADMMDeconvF2((), 50, rho_i, relu1) on the same input + chcat, as ONE
    grouped call (shared input, channel-concatenated output) == five separate layer calls concatenated."""
    d0 = dev()
    torch.manual_seed(3)
    y = torch.rand(2, 3, 256, 256, device=d0)
    rhos = [0.01, 0.03, 0.1, 0.3, 0.9]
    lam = torch.full((5,), 0.02, device=d0); rho = torch.tensor(rhos, device=d0)
    x = A.tvd_fft_grouped(y, lam, rho, None, iso, 50, groups=5, shared_input=True, channel_concat=True, activation="relu1")
    parts = [A.admm_layer_call(y, lam[g:g + 1].clone(), rho[g:g + 1].clone(), None, None, 50, iso, "relu1", 0.0, False, clamp=False)
             for g in range(5)]
    ref = torch.cat(parts, dim=1)
    assert x.shape == (2, 15, 256, 256)
    assert rel_l2(x.cpu(), ref.cpu()) < 1e-5
    xo = O.ACTIVATIONS["relu1"](O.tvd_fft_fast(A.to_julia(y.cpu()).double(), lam[2:3].cpu().double(), rho[2:3].cpu().double(), None, iso, 50))
    assert rel_l2(A.to_julia(x[:, 6:9].cpu().contiguous()), xo) < TOL


def test_fft_kernels_against_cufft_closed_form():
    """cuFFT (torch.fft on the GPU) as the oracle for the hand-written FFT kernels: after ONE iteration the
    solver's output is the closed form x_1 = F^-1( C conj(K) F y ), C = 1 / (|Sigma|^2 + rho |Lambda|^2)."""
    import math
    d0 = dev()
    torch.manual_seed(4)
    for (B, P, N, M, kh, kw) in [(2, 3, 512, 512, 15, 15), (1, 2, 2048, 128, 7, 10), (3, 1, 128, 1024, 4, 4), (2, 1, 4096, 32, 3, 3)]:
        y = torch.rand(B, P, N, M, device=d0)
        h = torch.rand(1, 1, kw, kh, device=d0) / (kh * kw)
        rho = torch.tensor([0.05], device=d0); lam = torch.tensor([0.01], device=d0)
        x = A.tvd_fft(y, lam, rho, h, False, 1)
        # cuFFT closed form in fp64 (axes: last = dim 1 (M), second-to-last = dim 2 (N))
        yd = y.double()
        hh = torch.zeros(N, M, dtype=torch.float64, device=d0)
        hh[:kw, :kh] = h[0, 0].double()
        Sig = torch.fft.fft2(hh)
        k1 = torch.arange(M, device=d0, dtype=torch.float64); k2 = torch.arange(N, device=d0, dtype=torch.float64).reshape(-1, 1)
        pd, pr = (kh - 1) // 2, (kw - 1) // 2
        K = Sig * torch.exp(2j * math.pi * (k1 * pd / M + k2 * pr / N))
        L = 4 * torch.sin(math.pi * k1 / M) ** 2 + 4 * torch.sin(math.pi * k2 / N) ** 2
        C = 1.0 / (Sig.abs() ** 2 + 0.05 * L)
        ref = torch.fft.ifft2(C * K.conj() * torch.fft.fft2(yd)).real
        assert rel_l2(x.cpu(), ref.cpu()) < 2e-6, (B, P, N, M)


@pytest.mark.parametrize("L", [96, 160, 192, 320, 384, 480, 640, 768, 960, 1280, 1536, 1920])
def test_mixed_radix_lengths_both_dims(L):
    """Non-power-of-two lengths (the reference's FFTW path takes any size; here 3- and 5-smooth ones)."""
    for (M, N) in ((L, 64), (96, L)):
        y, h, _ = make_case(M, N, 2, 1, 5, 5, L)
        x = run_gpu(y, h, 0.02, 0.1, False, 6)
        xo = oracle(y, h, 0.02, 0.1, False, 6, fast=True)
        assert rel_l2(x, xo) < TOL, (M, N)


def test_video_frame_sizes():
    """640x480 and 1280x960 RGB frames, 15x15 PSF, 30 iterations."""
    for (M, N) in ((640, 480), (1280, 960)):
        y, h, _ = make_case(M, N, 3, 1, 15, 15, M + N)
        x = run_gpu(y, h, 0.0041, 0.021, False, 30)
        xo = oracle(y, h, 0.0041, 0.021, False, 30, fast=True)
        assert rel_l2(x, xo) < TOL, (M, N)


def test_huge_plane_count_exceeds_grid_y_limit():
    """140001 independent 32x32 planes (70001 pairs > the 65535 gridDim.y limit; odd count -> padded last pair)."""
    d0 = dev()
    torch.manual_seed(9)
    B = 140001
    y = torch.rand(B, 1, 32, 32, device=d0)
    h = A.from_julia(O.gaussian_psf(5, 1.0).float()).to(d0)
    lam = torch.tensor([0.02], device=d0); rho = torch.tensor([0.1], device=d0)
    x = A.tvd_fft(y, lam, rho, h, False, 3)
    torch.cuda.synchronize()
    for sl in (slice(0, 3), slice(70000, 70003), slice(B - 3, B)):
        yj = A.to_julia(y[sl].cpu()).double()
        xo = O.tvd_fft_fast(yj, lam.cpu().double(), rho.cpu().double(), O.gaussian_psf(5, 1.0).float().double(), False, 3)
        assert rel_l2(A.to_julia(x[sl].cpu()), xo) < TOL


def test_full_cfg2_size_shift_equivariance_and_fixed_point():
    """Size-independent properties at the FULL BASELINE configs[1] size (64 x 512x512 RGB, 15x15 PSF): with circular
    boundaries the solver commutes with circular shifts of the input, and a constant image is a fixed point when
    the PSF sums to one."""
    d0 = dev()
    torch.manual_seed(11)
    B, P, N, M, K = 64, 3, 512, 512, 20
    h = A.from_julia(O.motion_psf(15, 0.7, 11.0).float()).to(d0)
    lam = torch.tensor([0.0041], device=d0); rho = torch.tensor([0.021], device=d0)
    y = torch.rand(B, P, N, M, device=d0)
    x = A.tvd_fft(y, lam, rho, h, False, K)
    xs = A.tvd_fft(torch.roll(y, shifts=(37, -5), dims=(2, 3)).contiguous(), lam, rho, h, False, K)
    ref = torch.roll(x, shifts=(37, -5), dims=(2, 3))
    err = float((xs - ref).norm() / ref.norm())
    assert err < 1e-5, err
    del xs, ref, x
    c = torch.full((B, P, N, M), 0.37, device=d0)
    xc = A.tvd_fft(c, lam, rho, h, False, K)
    assert float((xc - 0.37).abs().max()) < 2e-5


# ---- any image size (the reference's FFTW path takes every size, ops.jl:26,86): generic-size kernels ---------------
@pytest.mark.parametrize(
    "M,N,P,B,kh,kw,K,iso",
    [(20, 24, 1, 2, 3, 3, 4, False), (33, 17, 3, 1, 5, 4, 3, False), (7, 5, 1, 1, 0, 0, 3, False), (100, 100, 3, 2, 7, 7, 20, False),
     (321, 481, 3, 1, 9, 9, 10, False), (225, 64, 1, 2, 5, 5, 10, True), (127, 131, 1, 2, 5, 5, 10, False),   # prime lengths
     (64, 50, 1, 3, 0, 0, 5, True), (360, 640, 3, 1, 15, 15, 5, False),
     (360, 1280, 1, 2, 9, 9, 6, False), (720, 256, 3, 1, 7, 7, 6, True),      # M without a plan, N planned: generic dim-1 + tuned dim-2
     (512, 200, 1, 2, 7, 7, 6, False), (1280, 720, 1, 1, 9, 9, 4, True)],     # N without a plan, M planned: tuned dim-1 + generic dim-2
)
def test_forward_any_size_vs_oracle(M, N, P, B, kh, kw, K, iso):
    y, h, _ = make_case(M, N, P, B, kh, kw, 300 + M + N, psf="random")
    x = run_gpu(y, h, 0.0041, 0.021, iso, K)
    xo = oracle(y, h, 0.0041, 0.021, iso, K, fast=M * N > 40000)
    assert rel_l2(x.double(), xo) <= TOL


def test_hd_frame_1080x1920_properties():
    """A size the oracle would take minutes on: shift equivariance and the constant fixed point (size-independent)."""
    d0 = dev()
    torch.manual_seed(0)
    y = torch.nn.functional.avg_pool2d(torch.rand(1, 1, 1920, 1080, device=d0), 5, 1, 2)   # image-like, not white noise
    g1 = torch.exp(-0.5 * (torch.arange(5, device=d0) - 2.0) ** 2 / 1.5 ** 2)
    h = torch.outer(g1, g1).reshape(1, 1, 5, 5); h /= h.sum()
    lam = torch.tensor([0.0041], device=d0); rho = torch.tensor([0.021], device=d0)
    x = A.tvd_fft(y, lam, rho, h, False, 4)
    xs = A.tvd_fft(torch.roll(y, (37, 11), (2, 3)).contiguous(), lam, rho, h, False, 4)
    assert rel_l2(xs.cpu(), torch.roll(x, (37, 11), (2, 3)).cpu()) <= 2e-5    # two fp32 results, each within 1e-5 of exact
    c = A.tvd_fft(torch.full_like(y, 0.37), lam, rho, h, False, 4)
    assert float((c - 0.37).abs().max()) <= 1e-5


def test_edge_sizes_and_psf_as_large_as_the_image():
    """Smallest planes, an odd plane count (padded pair), a PSF covering the whole image, the longest generic length."""
    for (M, N, P, B, kh, kw, K, iso) in [(2, 2, 1, 1, 0, 0, 3, False), (3, 2, 1, 3, 3, 2, 3, True), (16, 12, 1, 1, 16, 12, 2, False),
                                         (32, 32, 1, 1, 32, 32, 2, False), (4095, 33, 1, 1, 3, 3, 2, False), (35, 4094, 1, 1, 3, 3, 2, False)]:
        y, h, _ = make_case(M, N, P, B, kh, kw, 400 + M + N, psf="random")
        if h is not None:
            h = h / h.sum()
        x = run_gpu(y, h, 0.0041, 0.021, iso, K)
        xo = oracle(y, h, 0.0041, 0.021, iso, K, fast=M * N > 40000)
        # a random PSF as large as the image has |Sigma| ~ 1/sqrt(MN) off DC: the division amplifies fp32 rounding (5e-5 there)
        assert rel_l2(x.double(), xo) <= (5e-5 if kh == M and kh > 0 else TOL), (M, N, kh, kw)


def test_grouped_per_image_generic_size():
    """BASELINE configs[4] semantics (per-image PSF / lambda / rho) at a size without a register-FFT plan."""
    d0 = dev()
    M, N, B, K = 50, 36, 4, 4
    ys, hs, refs, lams, rhos = [], [], [], [], []
    for b in range(B):
        y, h, _ = make_case(M, N, 1, 1, 3, 5, 700 + b)
        lams.append(0.01 * (b + 1)); rhos.append(0.05 * (b + 2))
        ys.append(y); hs.append(h)
        refs.append(oracle(y, h, lams[-1], rhos[-1], False, K))
    yt = A.from_julia(torch.cat(ys, dim=3).float()).to(d0)
    ht = torch.stack([A.from_julia(h.float()).reshape(1, 5, 3) for h in hs]).contiguous().to(d0)     # (G,1,kw,kh)
    x = A.tvd_fft_grouped(yt, torch.tensor(lams, device=d0), torch.tensor(rhos, device=d0), ht, False, K, groups=B)
    assert rel_l2(A.to_julia(x.cpu()).double(), torch.cat(refs, dim=3)) <= TOL


@pytest.mark.parametrize("iso", [False, True])
def test_forward_is_cuda_graph_capturable(iso):
    """The library only enqueues stream-ordered kernels and memsets, so a call can be captured once and replayed
    (serving small batches is launch-bound: tools/graph_bench.py).  Neither TV variant uses floating-point atomics
    (the isotropic per-pixel norm is summed over the plane pairs in a fixed order), so replay == eager bit for bit."""
    d0 = dev()
    y = torch.rand(2, 3, 64, 64, device=d0)
    h = torch.rand(1, 1, 5, 5, device=d0); h /= h.sum()
    lam = torch.tensor([0.0041], device=d0); rho = torch.tensor([0.021], device=d0)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        A.tvd_fft(y, lam, rho, h, iso, 6)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        xg = A.tvd_fft(y, lam, rho, h, iso, 6)
    y.copy_(torch.rand_like(y))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(xg, A.tvd_fft(y, lam, rho, h, iso, 6))     # no atomics on either path: bitwise reproducible


@pytest.mark.parametrize("M,P,B,kh,K", [(128, 1, 128, 9, 6), (128, 3, 64, 5, 4), (64, 1, 256, 0, 5), (32, 2, 200, 3, 7)])
def test_small_plane_persistent_kernel(M, P, B, kh, K):
    """k_small (kernels_small.cuh): >= 64 plane pairs of 32^2 / 64^2 / 128^2, anisotropic inference: every iteration inside one
    persistent kernel per pair.  Checked against the fp64 oracle on a sample of images, and against the two-launch path."""
    from admm_deconv_b200 import _lib
    d0 = dev()
    y, h, _ = make_case(M, M, P, B, kh, kh, 2200 + M)
    yt = A.from_julia(y.float()).to(d0)
    ht = None if h is None else A.from_julia(h.float()).to(d0)
    lam = torch.tensor([0.0041], device=d0); rho = torch.tensor([0.021], device=d0)
    x = A.tvd_fft(yt, lam, rho, ht, False, K)
    sel = [0, B // 3, B - 1]
    xo = oracle(y[..., sel], None if h is None else h, 0.0041, 0.021, False, K, fast=True)
    assert rel_l2(A.to_julia(x.cpu())[..., sel], xo) < TOL
    x2 = A.admm_layer_call(yt, lam, rho, ht, None, K, False, "identity", 0.0, False, clamp=False) if False else None
    # the general path on the same input (flag ADMMTV_FLAG_NO_SMALL through the raw ABI)
    import harness
    be = harness.GpuBackend(A.load())
    r = be.forward(y.float().numpy(), 0.0041, 0.021, None if h is None else h.float().numpy()[:, :, 0, 0], False, K, flags=1 | _lib.FLAG_NO_SMALL)
    assert rel_l2(A.to_julia(x.cpu()), torch.from_numpy(r["x"].get())) < 2e-6


def test_small_plane_kernel_per_image_psfs_and_activation():
    """BASELINE configs[4] shape in miniature: per-image PSFs and (lambda, rho), groups = B, relu1, through k_small."""
    d0 = dev()
    M, B, K, k = 128, 96, 5, 7
    rng = np.random.default_rng(4)
    ys, hs, lams, rhos, refs = [], [], [], [], []
    for b in range(B):
        lams.append(0.002 + 0.004 * float(rng.random())); rhos.append(0.02 + 0.05 * float(rng.random()))
    y, _, _ = make_case(M, M, 1, B, k, k, 31)
    hh = rng.random((B, k, k)); hh /= hh.sum(axis=(1, 2), keepdims=True)
    yt = A.from_julia(y.float()).to(d0)
    ht = torch.from_numpy(hh).float().reshape(B, 1, k, k).transpose(2, 3).contiguous().to(d0)      # (G,1,kw,kh)
    x = A.tvd_fft_grouped(yt, torch.tensor(lams, device=d0), torch.tensor(rhos, device=d0), ht, False, K, groups=B, activation="relu1")
    for b in (0, 40, B - 1):
        hb = torch.from_numpy(hh[b]).float().double().reshape(k, k, 1, 1)
        xo = O.ACTIVATIONS["relu1"](oracle(y[..., b:b + 1], hb, lams[b], rhos[b], False, K, fast=True))
        assert rel_l2(A.to_julia(x.cpu())[..., b:b + 1], xo) < TOL
