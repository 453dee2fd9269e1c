"""GPU test of the device-side batch assembly: ImageDataFeeder.getindex against a numpy restatement of
datafeeder.jl:31-68 + base_funcs.jl:29-35 (random aligned crops -> Float32 (H,W,C) -> cat on dim 4)."""
import numpy as np
import pytest
import torch

import admm_deconv_b200 as A
from admm_deconv_b200.staging import ImageDataFeeder

pytestmark = pytest.mark.gpu


def _dataset(n, H, W, C, seed):
    rng = np.random.default_rng(seed)
    shp = (H, W) if C == 1 else (H, W, C)
    xs = [rng.integers(0, 256, size=shp, dtype=np.uint8) for _ in range(n)]
    ys = [rng.integers(0, 256, size=shp, dtype=np.uint8) for _ in range(n)]
    return xs, ys


@pytest.mark.parametrize("C", [1, 3])
def test_feeder_matches_reference_assembly(C):
    xs, ys = _dataset(6, 300, 280, C, 3)
    f = ImageDataFeeder(xs, ys, (256, 256), (256, 256), "cuda:0", seed=11)
    idxs = [4, 0, 5, 2]
    origins = [(3, 7), (44, 0), (0, 24), (17, 17)]
    for rep in range(3):      # exercises the double buffering
        bx, by = f.getindex(idxs, origins)
        torch.cuda.synchronize()
        assert bx.shape == (4, C, 256, 256) and bx.dtype == torch.float32
        for data, got in ((xs, bx), (ys, by)):
            want = np.stack([data[i].reshape(300, 280, C)[h:h + 256, w:w + 256].astype(np.float32) / np.float32(255)
                             for i, (h, w) in zip(idxs, origins)], axis=-1)        # (H,W,C,B) as getindex builds it
            assert np.array_equal(A.to_julia(got.cpu()).numpy(), want)


def test_feeder_shards_and_feeds_the_layer():
    xs, ys = _dataset(8, 128, 128, 3, 5)
    f0 = ImageDataFeeder(xs, ys, (128, 128), (128, 128), "cuda:0", seed=1, rank=0, world=2)
    f1 = ImageDataFeeder(xs, ys, (128, 128), (128, 128), "cuda:0", seed=1, rank=1, world=2)
    full = ImageDataFeeder(xs, ys, (128, 128), (128, 128), "cuda:0", seed=1)
    idxs = [0, 1, 2, 3, 4]
    a, _ = f0.getindex(idxs); b, _ = f1.getindex(idxs); c, cy = full.getindex(idxs)
    assert a.shape[0] == 3 and b.shape[0] == 2
    assert torch.equal(torch.cat([a, b]), c)
    d = torch.device("cuda:0")
    x = A.tvd_fft(c, torch.tensor([0.0041], device=d), torch.tensor([0.021], device=d), None, False, 3)
    assert torch.isfinite(x).all() and float(A.gmsd_loss(x, cy)) > 0


def test_resident_dataset_gather_matches_streaming_feeder():
    """resident=True keeps the 8-bit dataset in HBM and gathers crops on the device: identical batches."""
    xs, ys = _dataset(5, 150, 140, 3, 9)
    a = ImageDataFeeder(xs, ys, (128, 96), (128, 96), "cuda:0", seed=2)
    b = ImageDataFeeder(xs, ys, (128, 96), (128, 96), "cuda:0", seed=2, resident=True)
    idxs = [3, 1, 4]
    origins = [(0, 0), (22, 44), (5, 17)]
    ax, ay = a.getindex(idxs, origins); bx, by = b.getindex(idxs, origins)
    torch.cuda.synchronize()
    assert torch.equal(ax, bx) and torch.equal(ay, by)
    assert b.h2d_bytes(3) == 48 and a.h2d_bytes(3) == 2 * 3 * 128 * 96 * 3


def test_feeder_back_to_back_without_synchronize():
    """Streaming feeder, several batches in flight with work queued on the compute stream and NO host synchronisation
    in between: every batch must still hold its own crops when it is consumed (the copy stream must not recycle a block
    that queued kernels are reading)."""
    xs, ys = _dataset(12, 140, 140, 3, 21)
    f = ImageDataFeeder(xs, ys, (128, 128), (128, 128), "cuda:0", seed=4)
    d = torch.device("cuda:0")
    sums, want = [], []
    burn = torch.rand(2048, 2048, device=d)
    for step in range(8):
        idxs = [(step * 3 + k) % 12 for k in range(4)]
        origins = [(step % 7, (2 * step) % 9)] * 4
        bx, by = f.getindex(idxs, origins)
        for _ in range(6):                      # keep the compute stream busy so the next getindex overlaps it
            burn = (burn @ burn).clamp_(-1, 1)
        sums.append((bx.double().sum() + 2 * by.double().sum()).reshape(1))      # consumed late, on the compute stream
        del bx, by                              # frees the blocks while their reads are still queued
        w = 0.0
        for data, m in ((xs, 1.0), (ys, 2.0)):
            for i, (h, ww) in zip(idxs, origins):
                w += m * float((data[i][h:h + 128, ww:ww + 128].astype(np.float32) / np.float32(255)).astype(np.float64).sum())
        want.append(w)
    got = torch.cat(sums).cpu().numpy()
    assert np.allclose(got, np.array(want), rtol=1e-9)
