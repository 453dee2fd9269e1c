"""CPU tests of the C-ABI boundary: the product library loads, exports every symbol
include/admmtv.h declares, and validates descriptors without a GPU.  No compute calls."""
import os
import re

import pytest

from admm_deconv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "admmtv.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(admmtv_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib.lib, name), name
    assert lib.version() == 100


def test_loss_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "admmtv_loss.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(admmtv_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.LOSS_SYMBOLS)
    for name in declared:
        assert hasattr(lib.lib, name), name
    # argument validation without a GPU
    assert lib.gmsd_workspace_bytes(64, 64, 3, 2) > 0
    assert lib.ssim_workspace_bytes(64, 64, 3, 2, None, True) >= 3 * 54 * 54 * 6 * 4
    with pytest.raises(_lib.AdmmTvError):
        lib.ssim_workspace_bytes(8, 8, 1, 1, None, True)


def test_host_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "admmtv_host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(admmtv_(?:host|mse)_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.HOST_SYMBOLS)
    for name in declared:
        assert hasattr(lib.lib, name), name
    # sizing and validation need no GPU
    d = _lib.make_desc(64, 64, 3, 2, 5, 5, 10)
    infer, train = lib.host_session_bytes(d, False), lib.host_session_bytes(d, True)
    fwd_b, ck_b, bwd_b = lib.workspace_bytes(d)
    assert infer >= fwd_b + 4 * 64 * 64 * 3 * 2 * 4 and train >= infer + ck_b + bwd_b
    assert lib.host_grad_floats(d) == 25 + 2
    with pytest.raises(_lib.AdmmTvError):
        lib.host_session_bytes(_lib.make_desc(64, 64, 3, 2, 5, 5, 0), True)


def test_batch_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "admmtv_batch.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(admmtv_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.BATCH_SYMBOLS)
    for name in declared:
        assert hasattr(lib.lib, name), name
    with pytest.raises(_lib.AdmmTvError):
        lib.batch_from_n0f8(0, 4, 3, 1, 0, 256, 1, 3, 12, 48, 256)   # invalid shape is rejected before any launch


def test_loss_cpu_tensor_is_rejected():
    import torch
    from admm_deconv_b200 import losses
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.gmsd(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))


def test_desc_layout_matches_header():
    import ctypes
    assert ctypes.sizeof(_lib.Desc) == 14 * 4


@pytest.mark.parametrize(
    "kw,code",
    [
        (dict(), 0),
        (dict(M=48), 0), (dict(N=16), 0), (dict(M=8192), -3), (dict(M=224), 0), (dict(M=96, N=1920), 0), (dict(M=481, N=321), 0),
        (dict(N=4097), -3),
        (dict(M=0), -2), (dict(kh=3, kw=0), -2), (dict(kh=65, kw=3, M=64), -2),
        (dict(iters=0), -4),
        (dict(iso=2), -5), (dict(activation=7), -5), (dict(flags=256), -5), (dict(flags=64 | 128), 0),
    ],
)
def test_check_error_codes(lib, kw, code):
    base = dict(M=64, N=64, P=3, B=2, kh=5, kw=5, iters=10, iso=0, activation=0, has_bias=0, device=0, flags=0, creg=0.0)
    base.update(kw)
    d = _lib.Desc(base["M"], base["N"], base["P"], base["B"], base["kh"], base["kw"], base["iters"], base["iso"],
                  base["activation"], base["has_bias"], base["device"], base["flags"], base["creg"], 0)
    assert lib.check(d) == code
    assert isinstance(lib.strerror(code), str) and len(lib.strerror(code)) > 0


def test_workspace_sizes_scale(lib):
    d = _lib.make_desc(512, 512, 3, 64, 15, 15, 100)
    fwd, ck, bwd = lib.workspace_bytes(d)
    px = 512 * 512 * 3 * 64
    assert 24 * px <= fwd <= 32 * px          # b 4 + spec 2x4 + v 2x8 B per plane-pixel (+ tables)
    assert ck >= 12 * px * 99


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(OSError):
        _lib.AdmmTvLib(str(tmp_path / "nope.so"))


def test_cpu_tensor_is_rejected():
    import torch
    from admm_deconv_b200 import ops
    y = torch.zeros(1, 1, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.tvd_fft(y, torch.tensor([0.1]), torch.tensor([0.1]), None, False, 2)
