"""Small end-to-end exercise of every kernel (aniso/iso, fwd/bwd, grouped, several FFT lengths) for
compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import admm_deconv_b200 as A  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (M, N, P, B, k) in [(32, 64, 3, 1, 5), (64, 32, 1, 3, 0), (128, 256, 2, 1, 7), (512, 32, 1, 2, 3), (32, 1024, 1, 1, 3), (2048, 32, 2, 1, 3)]:
    for iso in (False, True):
        y = torch.rand(B, P, N, M, device=dev, requires_grad=True)
        h = None if k == 0 else (torch.rand(1, 1, k, k, device=dev) / (k * k)).requires_grad_(True)
        lam = torch.tensor([0.02], device=dev, requires_grad=True)
        rho = torch.tensor([0.1], device=dev, requires_grad=True)
        bias = torch.tensor([0.01], device=dev, requires_grad=True)
        x = A.admm_layer_call(y, lam, rho, h, bias, 3, iso, "relu1", 0.0)
        x.sum().backward()
        torch.cuda.synchronize()
        print(M, N, P, B, k, iso, float(x.mean()), float(lam.grad), flush=True)
y = torch.rand(4, 1, 64, 64, device=dev)
h = torch.rand(4, 1, 5, 5, device=dev) / 25
x = A.tvd_fft_grouped(y, torch.full((4,), 0.02, device=dev), torch.full((4,), 0.1, device=dev), h, False, 3, groups=4)
x2 = A.tvd_fft_grouped(y[:2].contiguous(), torch.full((3,), 0.02, device=dev), torch.full((3,), 0.1, device=dev), None, True, 3, groups=3,
                       shared_input=True, channel_concat=True, activation="relu1")
torch.cuda.synchronize()
print("grouped ok", float(x.mean()), float(x2.mean()))
# sizes without a register-FFT plan (generic kernels, both mixed dispatches), losses, batch assembly
for (M, N, P, B, k, iso) in [(20, 24, 1, 2, 3, False), (33, 17, 3, 1, 5, True), (48, 32, 1, 2, 3, False), (32, 224, 1, 1, 3, True), (7, 5, 1, 1, 0, False)]:
    y = torch.rand(B, P, N, M, device=dev, requires_grad=True)
    h = None if k == 0 else (torch.rand(1, 1, k, k, device=dev) / (k * k)).requires_grad_(True)
    lam = torch.tensor([0.02], device=dev, requires_grad=True); rho = torch.tensor([0.1], device=dev, requires_grad=True)
    x = A.admm_layer_call(y, lam, rho, h, None, 3, iso, "relu1", 0.0)
    tgt = torch.rand_like(x)
    (A.gmsd_loss(x, tgt) + (A.ssim_loss(x, tgt) if min(M, N) >= 11 else 0.0)).backward()
    torch.cuda.synchronize()
    print("generic", M, N, P, B, k, iso, float(x.mean()), float(lam.grad), flush=True)
import numpy as np  # noqa: E402
from admm_deconv_b200.staging import ImageDataFeeder  # noqa: E402
rng = np.random.default_rng(0)
xs = [rng.integers(0, 256, size=(70, 50, 3), dtype=np.uint8) for _ in range(3)]
for resident in (False, True):
    f = ImageDataFeeder(xs, xs, (33, 40), (33, 40), dev, seed=0, resident=resident)
    bx, by = f.getindex([2, 0, 1])
    torch.cuda.synchronize()
    print("feeder", resident, float(bx.mean()))
# round 2: TMA-pipelined dim-2 pass (training forward at N = 512), per-iteration parameters, k_small, host-buffer session
y = torch.rand(2, 1, 512, 64, device=dev, requires_grad=True)
h = (torch.rand(1, 1, 3, 3, device=dev) / 9).requires_grad_(True)
lam = torch.full((3,), 0.02, device=dev, requires_grad=True); rho = torch.full((3,), 0.1, device=dev, requires_grad=True)
x = A.admm_layer_call(y, lam, rho, h, None, 3, False, "identity", 0.0)
x.sum().backward()
torch.cuda.synchronize()
print("tma + per-iteration ok", float(x.mean()), lam.grad.tolist(), flush=True)
for (M, Bs) in ((32, 130), (64, 128)):
    with torch.no_grad():
        ys = torch.rand(Bs, 1, M, M, device=dev)
        xs_ = A.tvd_fft(ys, torch.tensor([0.02], device=dev), torch.tensor([0.1], device=dev), torch.rand(1, 1, 3, 3, device=dev) / 9, False, 4)
    torch.cuda.synchronize()
    print("k_small ok", M, float(xs_.mean()), flush=True)
sess = A.host.HostSession(64, 32, 1, 2, 3, 3, iters=3, training=True)
yc = torch.rand(2, 1, 32, 64).pin_memory(); tc = torch.rand(2, 1, 32, 64).pin_memory()
g, l = sess.train_step(yc, tc, torch.tensor([0.02]), torch.tensor([0.1]), torch.rand(1, 1, 3, 3) / 9)
sess.close()
print("host session ok", float(l), flush=True)
