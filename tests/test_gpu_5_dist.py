"""GPU tests of the cross-rank isotropic coupling (admmtv_forward_ex / admmtv_backward_ex) on ONE GPU:
a world-size-1 NCCL group drives the per-iteration callback path; results must be bit-identical to the plain call.
The two-rank exactness check runs on CPU (tests/test_dist_gloo.py, emulated kernels + gloo) and on two GPUs with
tools/iso_coupling_check.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist

import admm_deconv_b200 as A
from admm_deconv_b200 import dist as D

pytestmark = pytest.mark.gpu


def test_iso_coupling_world1_is_identity():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
    dev = torch.device("cuda:0")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        torch.manual_seed(0)
        y = torch.rand(3, 2, 64, 64, device=dev)
        h = torch.rand(1, 1, 3, 3, device=dev) / 9
        outs = []
        xbar = torch.randn(3, 2, 64, 64, device=dev)      # (a constant cotangent has zero λ/ρ gradient: the mean of x is fixed)
        for cp in (None, D.IsoCoupling()):
            lam = torch.tensor([0.05], device=dev, requires_grad=True); rho = torch.tensor([0.3], device=dev, requires_grad=True)
            hh = h.clone().requires_grad_(True)
            x = A.admm_layer_call(y, lam, rho, hh, None, 5, True, "identity", 0.0, False, clamp=False, iso_coupling=cp)
            x.backward(xbar)
            torch.cuda.synchronize()
            outs.append((x.detach(), hh.grad, lam.grad, rho.grad))
            if cp is not None:
                assert cp.calls == 2 * 4          # K-1 forward + K-1 backward all-reduces
        # the per-pixel norm is a fixed-order sum (no float atomics): x is bit-identical with and without the callbacks;
        # the parameter gradients pass through fp64 atomics and are compared to rounding
        assert torch.equal(outs[0][0], outs[1][0])
        for a, b in zip(outs[0][1:], outs[1][1:]):
            assert float((a - b).norm() / b.norm()) < 1e-5
    finally:
        dist.destroy_process_group()


def test_iso_coupling_is_cuda_graph_capturable():
    """The per-iteration all-reduce hook is called while the library ENQUEUES (a host callback into torch.distributed).  Under
    CUDA-graph capture that happens once, at capture time: NCCL's kernels are recorded into the graph and a replay runs the
    whole coupled forward with no host round trip at all."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
    dev = torch.device("cuda:0")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        torch.manual_seed(1)
        y = torch.rand(3, 2, 64, 64, device=dev)
        h = torch.rand(1, 1, 3, 3, device=dev) / 9
        lam = torch.tensor([0.05], device=dev); rho = torch.tensor([0.3], device=dev)
        cp = D.IsoCoupling()
        run = lambda: A.admm_layer_call(y, lam, rho, h, None, 5, True, "identity", 0.0, False, clamp=False, iso_coupling=cp)
        dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()       # communicator up before the capture
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st), torch.no_grad():
            run()
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        calls = cp.calls
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            xg = run()
        assert cp.calls == calls + 4            # K-1 hook calls, made once while capturing
        y.copy_(torch.rand_like(y))
        g.replay(); g.replay()
        torch.cuda.synchronize()
        assert cp.calls == calls + 4            # replays do not come back to the host
        with torch.no_grad():
            assert torch.equal(xg, run())
    finally:
        dist.destroy_process_group()
