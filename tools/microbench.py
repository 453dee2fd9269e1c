"""Times one workload's forward (and optionally fwd+bwd) for several library variants.
    python tools/microbench.py cfg2 20 base swz swz_pf ...        (tags of tools/variants.py; 'main' = libadmmtv.so)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from admm_deconv_b200 import _lib, ops  # noqa: E402

name, iters = sys.argv[1], int(sys.argv[2])
tags = [a for a in sys.argv[3:] if not a.startswith("--")]
ISO = "--iso" in sys.argv
NOPSF = "--nopsf" in sys.argv
w = dict(bench.WORKLOADS[name], iters=iters)
y, _, h = bench.make_inputs(w, 1001)
dev = torch.device("cuda:0")
y = y.to(dev); h = None if NOPSF else h.to(dev)
hp = None if h is None else h.data_ptr()
px = y.numel()
ref = None
for tag in tags:
    path = _lib.LIB_PATH if tag == "main" else os.path.join(ROOT, "admm_deconv_b200", f"libadmmtv_{tag}.so")
    lib = _lib.AdmmTvLib(path)
    d = ops.make_desc_for(y, h, iters, ISO, "identity", False, _lib.FLAG_NO_CLAMP, 0.0)
    fwd_b, ck_b, bwd_b = lib.workspace_bytes(d)
    ws = torch.empty(fwd_b, dtype=torch.uint8, device=dev)
    x = torch.empty_like(y)
    lam = torch.tensor([0.0041], device=dev); rho = torch.tensor([0.021], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    args = (d, y.data_ptr(), hp, lam.data_ptr(), rho.data_ptr(), None, x.data_ptr(), ws.data_ptr(), None, st)
    for _ in range(2):
        lib.profile_forward(*args)
    res = [lib.profile_forward(*args) for _ in range(3)]
    tot, t2, t1, oth = [min(r[i] for r in res) for i in range(4)]
    if ref is None:
        ref = x.clone()
    err = float((x - ref).norm() / ref.norm())
    it = t2 / iters + t1 / max(iters - 1, 1)
    # forward with checkpoint + backward through the raw ABI
    ck = torch.empty(ck_b, dtype=torch.uint8, device=dev)
    wsb = torch.empty(bwd_b, dtype=torch.uint8, device=dev)
    xbar = torch.ones_like(y); ybar = torch.empty_like(y); hbar = None if h is None else torch.empty_like(h)
    lb = torch.empty(1, device=dev); rb = torch.empty(1, device=dev)
    def train():
        lib.forward(d, y.data_ptr(), hp, lam.data_ptr(), rho.data_ptr(), None, x.data_ptr(), ws.data_ptr(), ck.data_ptr(), st)
        lib.backward(d, xbar.data_ptr(), x.data_ptr(), y.data_ptr(), hp, lam.data_ptr(), rho.data_ptr(), ck.data_ptr(),
                     ybar.data_ptr(), (None if hbar is None else hbar.data_ptr()), lb.data_ptr(), rb.data_ptr(), None, wsb.data_ptr(), st)
    train(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); train(); train(); e1.record(); torch.cuda.synchronize()
    tr = e0.elapsed_time(e1) / 2
    pf = lib.profile_forward(d, y.data_ptr(), hp, lam.data_ptr(), rho.data_ptr(), None, x.data_ptr(), ws.data_ptr(), ck.data_ptr(), st)
    pb = lib.profile_backward(d, xbar.data_ptr(), x.data_ptr(), y.data_ptr(), hp, lam.data_ptr(), rho.data_ptr(), ck.data_ptr(),
                              ybar.data_ptr(), (None if hbar is None else hbar.data_ptr()), lb.data_ptr(), rb.data_ptr(), None, wsb.data_ptr(), st)
    print(f"{tag:12s} ckpt-fwd: dim2 {pf[1]/iters*1e3:6.1f} dim1 {pf[2]/max(iters-1,1)*1e3:6.1f} other {pf[3]:.3f} ms | bwd: dim2 {pb[1]/iters*1e3:6.1f} "
          f"dim1 {pb[2]/max(iters-1,1)*1e3:6.1f} other {pb[3]:.3f} ms total {pb[0]:.2f} ms")
    print(f"{tag:12s} train {tr:7.2f} ms ({108*px*iters/tr/1e6/6538.6:.3f})  dim2 {t2/iters*1e3:7.1f} us  dim1 {t1/max(iters-1,1)*1e3:7.1f} us  iter {it*1e3:7.1f} us  "
          f"alg {40*px/it/1e6:7.0f} GB/s ({40*px/it/1e6/6538.6:.3f})  other {oth:.3f} ms  diff_vs_first {err:.1e}", flush=True)
