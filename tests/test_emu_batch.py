"""CPU test of the batch-assembly kernel (include/admmtv_batch.h) through the emulated build."""
import numpy as np
import pytest

import emu_harness as E


@pytest.mark.parametrize("M,N,C,B", [(40, 33, 3, 2), (32, 64, 1, 3), (7, 5, 5, 1)])
@pytest.mark.parametrize("order", ["numpy", "julia", "planar"])
def test_batch_from_n0f8(emu, M, N, C, B, order):
    rng = np.random.default_rng(M * N + C)
    img = rng.integers(0, 256, size=(B, M, N, C), dtype=np.uint8)     # logical (b, i, j, c)
    if order == "numpy":       # row-major (H,W,C) crops
        src = np.ascontiguousarray(img)
        sc, si, sj, sb = 1, C * N, C, M * N * C
    elif order == "julia":     # Julia Matrix{RGB{N0f8}} crops: c fastest, then i, then j (vectorised path when M N % 4 == 0)
        src = np.ascontiguousarray(img.transpose(0, 2, 1, 3))
        sc, si, sj, sb = 1, C, C * M, M * N * C
    else:                      # planar bytes in the destination's own (M,N,C,B) order: the flat vectorised path (+ its tail)
        src = np.ascontiguousarray(img.transpose(0, 3, 2, 1))
        sc, si, sj, sb = M * N, 1, M, M * N * C
    dst = np.asfortranarray(np.full((M, N, C, B), np.nan, np.float32))
    emu.batch_from_n0f8(M, N, C, B, 0, src.ctypes.data, sc, si, sj, sb, dst.ctypes.data)
    want = (img.transpose(1, 2, 3, 0).astype(np.float32) / np.float32(255.0))   # img2tensor: N0f8 -> Float32
    assert np.array_equal(np.array(dst), want)                         # bit-exact: integer -> correctly rounded i/255


def test_batch_gather_from_resident_images(emu):
    rng = np.random.default_rng(4)
    H, W, C, B, M, N = 50, 44, 3, 3, 33, 20
    imgs = rng.integers(0, 256, size=(4, H, W, C), dtype=np.uint8)            # row-major (H,W,C) images, back to back
    pick, org = [2, 0, 3], [(0, 0), (17, 24), (5, 9)]
    offs = np.array([i * H * W * C + (h0 * W + w0) * C for i, (h0, w0) in zip(pick, org)], dtype=np.int64)
    dst = np.asfortranarray(np.full((M, N, C, B), np.nan, np.float32))
    emu.batch_gather_n0f8(M, N, C, B, 0, imgs.ctypes.data, offs.ctypes.data, 1, C * W, C, dst.ctypes.data)
    want = np.stack([imgs[i, h0:h0 + M, w0:w0 + N].astype(np.float32) / np.float32(255) for i, (h0, w0) in zip(pick, org)], axis=-1)
    assert np.array_equal(np.array(dst), want)
