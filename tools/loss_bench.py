"""Times the loss kernels (include/admmtv_loss.h) on the GPU and reports achieved HBM GB/s against the measured
copy peak (MEASURED_PEAKS.json).  Algorithmic bytes per pixel: GMSD forward 8 (x, y read), backward 12 (x, y read,
xbar written); SSIM forward 8 (+12 per output pixel for the derivative maps when training), backward 24.
    python tools/loss_bench.py [B C N M]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import admm_deconv_b200 as A  # noqa: E402

B, C, N, M = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (64, 3, 512, 512)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6538.6
dev = torch.device("cuda:0")
x = torch.rand(B, C, N, M, device=dev, requires_grad=True)
y = (x.detach() + 0.05 * torch.randn_like(x)).clamp(0, 1)
px = x.numel()


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
for name, fn, bf, bb in (("gmsd_loss", A.gmsd_loss, 8, 12), ("ssim_loss", A.ssim_loss, 20, 24)):
    with torch.no_grad():
        tf_inf = timeit(lambda: fn(x, y))
    loss = fn(x, y)
    tf = timeit(lambda: fn(x, y))
    tb = timeit(lambda: torch.autograd.grad(loss, x, retain_graph=True))
    inf_bytes = 8
    res[name] = dict(fwd_inference_ms=tf_inf, fwd_inference_GBs=inf_bytes * px / tf_inf / 1e6, fwd_train_ms=tf,
                     fwd_train_GBs=bf * px / tf / 1e6, bwd_ms=tb, bwd_GBs=bb * px / tb / 1e6,
                     fwd_bwd_frac_of_peak=(bf + bb) * px / (tf + tb) / 1e6 / PEAK)
print(json.dumps({"shape": [B, C, N, M], "peak_GBs": PEAK, **res}))
