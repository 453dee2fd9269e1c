"""CPU tests of the kernel SOURCES through the thread-per-CUDA-thread emulation (tests/emu/):
same .cu/.cuh files, same C ABI, numpy buffers instead of device memory.  This checks the index
arithmetic / halo / barrier logic of the kernels where no GPU exists; the parity tests proper are
the -m gpu tests, which call the nvcc-built library."""
import glob
import os

import numpy as np
import pytest
import torch

import emu_harness as E
from cases import make_case, rel_l2
from oracle import admm_tv_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-5   # north_star: relative L2 <= 1e-5 (fp32) on the restored image


def _run(emu, y, h, lam, rho, iso, K, **kw):
    r = E.forward(emu, y.numpy(), lam, rho, None if h is None else h.numpy()[:, :, 0, 0], iso, K, **kw)
    return torch.from_numpy(np.ascontiguousarray(r["x"])), r


@pytest.mark.parametrize(
    "M,N,P,B,kh,kw,K",
    [(32, 32, 1, 2, 0, 0, 1), (32, 32, 1, 1, 0, 0, 4), (32, 64, 3, 1, 5, 4, 3), (64, 32, 1, 3, 3, 3, 5), (128, 32, 1, 2, 7, 7, 3)],
)
def test_emu_forward_aniso_vs_oracle(emu, M, N, P, B, kh, kw, K):
    y, h, _ = make_case(M, N, P, B, kh, kw, 100 + M + N + K)
    y = y.float().double()
    h = None if h is None else h.float().double()
    lam, rho = 0.05, 0.3
    x, _ = _run(emu, y, h, lam, rho, False, K, flags=1)
    lam_t = torch.tensor([lam], dtype=torch.float32).double()
    rho_t = torch.tensor([rho], dtype=torch.float32).double()
    xo = O.tvd_fft_cpu(y, lam_t, rho_t, h, False, K)
    assert rel_l2(x, xo) < TOL


@pytest.mark.parametrize("L", [256, 512, 1024, 2048, 4096])
def test_emu_every_fft_length_both_dims(emu, L):
    for (M, N) in ((L, 32), (32, L)):
        y, h, _ = make_case(M, N, 2, 1, 3, 3, L)
        y = y.float().double(); h = h.float().double()
        x, _ = _run(emu, y, h, 0.05, 0.3, False, 2, flags=1)
        xo = O.tvd_fft_fast(y, torch.tensor([0.05], dtype=torch.float32).double(),
                            torch.tensor([0.3], dtype=torch.float32).double(), h, False, 2)
        assert rel_l2(x, xo) < TOL, (M, N)


def test_emu_golden_forward(emu):
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "aniso_*.npz"))):
        d = np.load(f)
        if d["y"].shape[0] > 64:
            continue
        h = d["h"][:, :, 0, 0] if "h" in d else None
        r = E.forward(emu, d["y"], float(d["lam"]), float(d["rho"]), h, bool(d["iso"]), int(d["iters"]),
                      act=str(d["act"]), bias=float(d["bias"]) if "bias" in d else None, creg=float(d["creg"]))
        assert rel_l2(torch.from_numpy(np.ascontiguousarray(r["x"])), torch.from_numpy(d["x"])) < TOL, f


def test_emu_clamp_is_persisted(emu):
    y, h, _ = make_case(32, 32, 1, 2, 3, 3, 4)
    hh = h.numpy()[:, :, 0, 0].copy()
    hh[0, 0] = -0.5; hh[1, 1] = 1.7
    r = E.forward(emu, y.numpy(), -1.0, 0.01, hh, False, 2, creg=0.05)
    assert r["lam"][0] == np.float32(0.05) and r["rho"][0] == np.float32(0.05)     # deconv_admm.jl:216-217
    assert r["h"][0, 0] == 0.0 and r["h"][1, 1] == 1.0                              # :219
    # Julia's clamp(NaN, lo, hi) is NaN: a diverged parameter stays visible instead of being reset to the bound
    r = E.forward(emu, y.numpy(), float("nan"), 0.01, hh, False, 2, creg=0.05)
    assert np.isnan(r["lam"][0]) and r["rho"][0] == np.float32(0.05)


@pytest.mark.parametrize("M,N,P,B,kh,kw,K", [(32, 32, 1, 2, 0, 0, 2), (32, 64, 3, 1, 5, 4, 4), (64, 32, 1, 3, 3, 3, 5)])
def test_emu_forward_iso_vs_oracle(emu, M, N, P, B, kh, kw, K):
    """Isotropic (BT, ops.jl:10): ONE norm per pixel over both directions, all channels, all images."""
    import harness
    from parity import check_forward
    y, h, _ = make_case(M, N, P, B, kh, kw, 50 + M + K)
    check_forward(harness.EmuBackend(emu), y, h, 0.05, 0.3, True, K, fast=False)


def test_emu_golden_forward_iso(emu):
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "iso_*.npz"))):
        d = np.load(f)
        h = d["h"][:, :, 0, 0] if "h" in d else None
        r = E.forward(emu, d["y"], float(d["lam"]), float(d["rho"]), h, True, int(d["iters"]), act=str(d["act"]),
                      bias=float(d["bias"]) if "bias" in d else None, creg=float(d["creg"]))
        assert rel_l2(torch.from_numpy(np.ascontiguousarray(r["x"])), torch.from_numpy(d["x"])) < TOL, f


@pytest.mark.parametrize("M,N", [(96, 32), (32, 160), (192, 96), (320, 32), (32, 384), (480, 32), (32, 640), (768, 32)])
def test_emu_mixed_radix_lengths(emu, M, N):
    """3- and 5-smooth lengths (radix-3 / radix-5 first passes): 96, 160, 192, 320, 384, 480, 640, 768, ..."""
    import harness
    from parity import check_forward
    y, h, _ = make_case(M, N, 1, 2, 3, 3, 50 + M + N)
    check_forward(harness.EmuBackend(emu), y, h, 0.05, 0.3, False, 3)


# ---- sizes without a register-FFT plan: the generic kernels (generic_kernels.cuh), same C ABI -----------------------
@pytest.mark.parametrize(
    "M,N,P,B,kh,kw,K,iso",
    [(20, 24, 1, 2, 3, 3, 4, False), (33, 17, 3, 1, 5, 4, 3, False), (7, 5, 1, 1, 0, 0, 3, False), (31, 64, 1, 2, 3, 3, 3, False),
     (20, 24, 2, 2, 3, 3, 4, True), (64, 50, 1, 3, 0, 0, 3, True), (3, 9, 1, 1, 0, 0, 2, False),
     (48, 32, 1, 2, 3, 3, 3, False), (48, 64, 3, 1, 5, 4, 3, True),     # M generic, N planned: generic dim-1 + tuned dim-2
     (32, 224, 1, 2, 3, 3, 3, False), (64, 96 + 128, 1, 1, 0, 0, 3, True)],  # M planned, N = 224 = 7*32: tuned dim-1 + generic dim-2
)
def test_emu_forward_generic_sizes(emu, M, N, P, B, kh, kw, K, iso):
    import harness
    from cases import make_case, rel_l2
    from oracle import admm_tv_oracle as O
    be = harness.EmuBackend(emu)
    y, h, _ = make_case(M, N, P, B, kh, kw, 5 + M)
    f = be.forward(y.numpy(), 0.05, 0.3, None if h is None else h.numpy()[:, :, 0, 0], iso, K)
    x = torch.from_numpy(f["x"].get()).double()
    l = torch.tensor([0.05], dtype=torch.float32).double(); r = torch.tensor([0.3], dtype=torch.float32).double()
    xo = O.tvd_fft_cpu(y.float().double(), l, r, None if h is None else h.float().double(), iso, K)
    assert rel_l2(x, xo) <= 1e-5


@pytest.mark.parametrize("M,P,B,kh,K,per_image", [(32, 1, 128, 3, 5, False), (32, 2, 64, 0, 4, False), (64, 1, 64, 5, 3, True), (32, 1, 65, 3, 4, True)])
def test_emu_small_plane_persistent_kernel(emu, M, P, B, kh, K, per_image):
    """k_small (kernels_small.cuh): planes of 32^2 / 64^2 / 128^2 with >= 64 plane pairs run all iterations inside one
    persistent kernel per pair (the two-launch path agrees to rounding only -- different pass order -- so both are compared
    with the fp64 oracle).  per_image: groups = B single-plane images with their own PSF / lambda / rho: images 2q, 2q+1
    share one complex transform (mirrored spectral division); an odd B leaves the last pair half empty."""
    import numpy as np
    import harness
    from admm_deconv_b200 import _lib
    from oracle import admm_tv_oracle as O
    be = harness.EmuBackend(emu)
    y, h, _ = make_case(M, M, P, B, kh, kh, 2100 + M)
    if not per_image:
        for flags in (1, 1 | _lib.FLAG_NO_SMALL):
            r = be.forward(y.numpy(), 0.0041, 0.021, None if h is None else h.numpy()[:, :, 0, 0], False, K, flags=flags)
            xo = O.tvd_fft_fast(y.float().double(), torch.tensor([0.0041]).float().double(), torch.tensor([0.021]).float().double(),
                                None if h is None else h.float().double(), False, K)
            assert rel_l2(torch.from_numpy(r["x"].get()), xo) < 1e-5
        return
    rng = np.random.default_rng(3)
    hs = rng.random((kh, kh, B)); hs /= hs.sum(axis=(0, 1), keepdims=True)
    lams = 0.002 + 0.004 * rng.random(B); rhos = 0.02 + 0.05 * rng.random(B)
    x = be.forward_grouped(y.numpy(), lams, rhos, hs, False, K, groups=B)
    for b in (0, 17, B - 1):
        xo = O.tvd_fft_fast(y[..., b:b + 1].float().double(), torch.tensor([lams[b]]).float().double(), torch.tensor([rhos[b]]).float().double(),
                            torch.from_numpy(hs[:, :, b]).float().double().reshape(kh, kh, 1, 1), False, K)
        assert rel_l2(torch.from_numpy(x[..., b:b + 1]), xo) < 1e-5
