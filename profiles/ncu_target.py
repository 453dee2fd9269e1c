"""Short single-GPU target for ncu: cfg2 shapes (64 x 512x512 RGB, 15x15 PSF), a few iterations,
two forward calls (first = warm-up).  Usage: python profiles/ncu_target.py [workload] [iters] [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import admm_deconv_b200 as A  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = sys.argv[3] if len(sys.argv) > 3 else "fwd"
iso = len(sys.argv) > 4 and sys.argv[4] == "iso"
nopsf = len(sys.argv) > 5 and sys.argv[5] == "nopsf"
w = dict(bench.WORKLOADS[name], iters=iters)
y, _, h = bench.make_inputs(w, 1001)
dev = torch.device("cuda:0")
y = y.to(dev); h = None if nopsf else h.to(dev)
lam = torch.tensor([0.0041], device=dev); rho = torch.tensor([0.021], device=dev)
for rep in range(2):
    if mode == "fwd":
        x = A.tvd_fft(y, lam, rho, h, iso, iters)
    else:
        l = lam.clone().requires_grad_(True); r = rho.clone().requires_grad_(True); hh = None if h is None else h.clone().requires_grad_(True)
        x = A.admm_layer_call(y, l, r, hh, None, iters, iso, "identity", 0.0, False, clamp=False)
        x.backward(torch.ones_like(x))
    torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
