"""CPU test of the batch-assembly kernel (include/admmtv_batch.h) through the emulated build."""
import numpy as np
import pytest

import emu_harness as E


@pytest.mark.parametrize("M,N,C,B", [(40, 33, 3, 2), (32, 64, 1, 3), (7, 5, 5, 1)])
@pytest.mark.parametrize("order", ["numpy", "julia"])
def test_batch_from_n0f8(emu, M, N, C, B, order):
    rng = np.random.default_rng(M * N + C)
    img = rng.integers(0, 256, size=(B, M, N, C), dtype=np.uint8)     # logical (b, i, j, c)
    if order == "numpy":       # row-major (H,W,C) crops
        src = np.ascontiguousarray(img)
        sc, si, sj, sb = 1, C * N, C, M * N * C
    else:                      # Julia Matrix{RGB{N0f8}} crops: c fastest, then i, then j
        src = np.ascontiguousarray(img.transpose(0, 2, 1, 3))
        sc, si, sj, sb = 1, C, C * M, M * N * C
    dst = np.asfortranarray(np.full((M, N, C, B), np.nan, np.float32))
    emu.batch_from_n0f8(M, N, C, B, 0, src.ctypes.data, sc, si, sj, sb, dst.ctypes.data)
    want = (img.transpose(1, 2, 3, 0).astype(np.float32) / np.float32(255.0))   # img2tensor: N0f8 -> Float32
    assert np.array_equal(np.array(dst), want)                         # bit-exact: integer -> correctly rounded i/255
