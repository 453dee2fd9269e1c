"""The reference's shipped training configuration (configs/train_cfg.json: batch 2, 256x256 crops, use_iso=true;
net_build.jl:113-128: five ADMMDeconvF2((), 50, rho_i, relu1) branches on the same input, chcat): forward+backward
of the branch bank as ONE grouped call (ADMMParallel) versus five separate layer calls."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import admm_deconv_b200 as A  # noqa: E402

d = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rhos = [0.002, 0.02, 0.2, 2.0, 4.0]
layers = [A.ADMMDeconvF2((), 50, r, "relu1", iso=True).to(d) for r in rhos]
for l in layers:
    with torch.no_grad():
        l.lam.fill_(0.02)
bank = A.ADMMParallel(*layers)
x = torch.rand(B, 3, 256, 256, device=d)
tgt = torch.rand(B, 15, 256, 256, device=d)


def fused():
    for l in layers:
        l.lam.grad = None
    ((bank(x) - tgt) ** 2).mean().backward()


def separate():
    for l in layers:
        l.lam.grad = None
    ((torch.cat([l(x) for l in layers], dim=1) - tgt) ** 2).mean().backward()


for name, fn in (("fused", fused), ("separate", separate)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    px = B * 15 * 256 * 256 * 50
    print(f"{name:9s} batch {B}: {ms:8.3f} ms / training step of the 5-branch bank  ({px / ms / 1e3:9.0f} Mpx-it/s)", flush=True)
