"""CUDA-graph capture of the forward call (serving small batches is launch-bound: ~2K+6 launches of a few microseconds).
    python tools/graph_bench.py [B P N M K]
The library only enqueues stream-ordered work (kernels + memsets), so a call can be captured once and replayed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import admm_deconv_b200 as A  # noqa: E402

B, P, N, M, K = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (4, 1, 128, 128, 50)
dev = torch.device("cuda:0")
y = torch.rand(B, P, N, M, device=dev)
h = torch.rand(1, 1, 9, 9, device=dev); h /= h.sum()
lam = torch.tensor([0.0041], device=dev); rho = torch.tensor([0.021], device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


x_eager = A.tvd_fft(y, lam, rho, h, False, K)
t_eager = timeit(lambda: A.tvd_fft(y, lam, rho, h, False, K))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    A.tvd_fft(y, lam, rho, h, False, K)          # warm-up on the capture stream
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    x_graph = A.tvd_fft(y, lam, rho, h, False, K)
g.replay(); torch.cuda.synchronize()
t_graph = timeit(g.replay)
y.copy_(torch.rand_like(y)); g.replay(); torch.cuda.synchronize()     # new input through the same graph
ok = float((x_graph - A.tvd_fft(y, lam, rho, h, False, K)).abs().max())
px = B * P * N * M
print(json.dumps({"workload": f"{B} x {M}x{N}x{P}, 9x9 PSF, {K} iterations forward", "eager_ms": t_eager, "graph_ms": t_graph,
                  "speedup": t_eager / t_graph, "Mpx_it_per_s_graph": px * K / t_graph / 1e3, "max_abs_diff_replay_vs_eager": ok}))
