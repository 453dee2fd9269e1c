"""Two-GPU check of the exact global-batch isotropic TV, launched with torchrun:
    gpurun --gpus 2 -- python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/iso_coupling_check.py
Every rank runs its shard of the batch with dist.IsoCoupling (NCCL all-reduce of the per-pixel partial sums, once per
iteration); rank 0 also runs the WHOLE batch on its own GPU through the same library and compares."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import admm_deconv_b200 as A  # noqa: E402
from admm_deconv_b200 import dist as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
B, P, N, M, K = (int(os.environ.get("ISO_B", "8")), 3, 256, 256, int(os.environ.get("ISO_K", "20")))
gen = torch.Generator().manual_seed(3)
y = torch.rand(B, P, N, M, generator=gen).to(dev)
xbar = torch.randn(B, P, N, M, generator=gen).to(dev)
h = (torch.rand(1, 1, 5, 5, generator=gen) / 25).to(dev)


def run(ysh, xb, cp):
    lam = torch.tensor([0.05], device=dev, requires_grad=True); rho = torch.tensor([0.3], device=dev, requires_grad=True)
    hh = h.clone().requires_grad_(True)
    ysh = ysh.clone().requires_grad_(True)
    x = A.admm_layer_call(ysh, lam, rho, hh, None, K, True, "identity", 0.0, False, clamp=False, iso_coupling=cp)
    x.backward(xb)
    return x.detach(), ysh.grad, torch.cat([hh.grad.reshape(-1), lam.grad, rho.grad])


lo, hi = D.shard_range(B, rank, world)
cp = D.IsoCoupling()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run(y[lo:hi].contiguous(), xbar[lo:hi].contiguous(), cp)
torch.cuda.synchronize(); dist.barrier()
e0.record()
xs, ybs, pk = run(y[lo:hi].contiguous(), xbar[lo:hi].contiguous(), cp)
dist.all_reduce(pk)
e1.record(); torch.cuda.synchronize()
xs_all = [torch.empty_like(xs) for _ in range(world)]; yb_all = [torch.empty_like(ybs) for _ in range(world)]
dist.all_gather(xs_all, xs); dist.all_gather(yb_all, ybs)
# timing at scale: forward + backward of this rank's share, uncoupled vs coupled (eager: one host callback per iteration and
# direction) vs the coupled forward replayed from a CUDA graph (the callbacks ran once, at capture)
def timed(fn, n=10):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
ysh, xsh = y[lo:hi].contiguous(), xbar[lo:hi].contiguous()
t_plain = timed(lambda: run(ysh, xsh, None))
t_coupled = timed(lambda: run(ysh, xsh, cp))
lam0 = torch.tensor([0.05], device=dev); rho0 = torch.tensor([0.3], device=dev)
fwd = lambda c: A.admm_layer_call(ysh, lam0, rho0, h, None, K, True, "identity", 0.0, False, clamp=False, iso_coupling=c)
with torch.no_grad():
    t_fwd_plain = timed(lambda: fwd(None)); t_fwd_coupled = timed(lambda: fwd(cp))
    st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fwd(cp)
    torch.cuda.current_stream().wait_stream(st); torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        xg = fwd(cp)
    t_fwd_graph = timed(g.replay)
if rank == 0:
    xf, ybf, pkf = run(y, xbar, None)
    xs_shard, _, _ = run(y[lo:hi].contiguous(), xbar[lo:hi].contiguous(), None)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    print({"x_rel": rel(torch.cat(xs_all), xf), "ybar_rel": rel(torch.cat(yb_all), ybf), "grads_rel": rel(pk, pkf),
           "x_rel_without_coupling": rel(xs_shard, xf[lo:hi]), "allreduce_calls": cp.calls,
           "ms_fwd_bwd_coupled_first": e0.elapsed_time(e1), "world": world, "images_per_rank": hi - lo, "K": K,
           "ms_fwd_bwd_uncoupled": t_plain, "ms_fwd_bwd_coupled": t_coupled, "ms_fwd_uncoupled": t_fwd_plain,
           "ms_fwd_coupled_eager": t_fwd_coupled, "ms_fwd_coupled_cuda_graph": t_fwd_graph})
dist.destroy_process_group()
