"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, gradient packing and
the single all-reduce.  The per-rank arithmetic here is the ORACLE (no GPU in this container); what
is checked is the sharding semantics of SURVEY.md 8e: for anisotropic TV, forward results of the
shards concatenate to the full-batch result and the summed shard gradients equal the full-batch
parameter gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from admm_deconv_b200 import dist as D
from admm_deconv_b200.layers import ADMMDeconv
from cases import make_case, rel_l2
from oracle import admm_tv_oracle as O


def test_shard_range_covers_batch_contiguously():
    for B in (1, 2, 7, 64, 1024):
        for world in (1, 2, 4, 8):
            got = [D.shard_range(B, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == B
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        M, N, P, B, K = 16, 16, 3, 4, 5
        y, h, g = make_case(M, N, P, B, 5, 5, 31)
        lam = torch.tensor([0.03], dtype=torch.float64)
        rho = torch.tensor([0.4], dtype=torch.float64)
        xbar = 2.0 * (y - g)
        # this rank's shard: contiguous block of whole images (Julia dim 4)
        lo, hi = D.shard_range(B, rank, world)
        _, gr = O.layer_grads(y[..., lo:hi], xbar[..., lo:hi], h, None, lam, rho, K, False)
        # a layer object only as the parameter container the all-reduce helper works on
        layer = ADMMDeconv((5, 5), K)
        layer.weight.grad = gr["weight"].permute(3, 2, 1, 0).float().contiguous()
        layer.lam.grad = gr["lam"].float()
        layer.rho.grad = gr["rho"].float()
        n = D.allreduce_layer_grads(layer)
        assert n == 25 + 2
        if rank == 0:
            out["weight"] = layer.weight.grad.clone()
            out["lam"] = layer.lam.grad.clone()
            out["rho"] = layer.rho.grad.clone()
            xs = O.tvd_fft_cpu(y[..., lo:hi], lam, rho, h, False, K)
            out["x0"] = xs
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    M, N, P, B, K = 16, 16, 3, 4, 5
    y, h, g = make_case(M, N, P, B, 5, 5, 31)
    lam = torch.tensor([0.03], dtype=torch.float64)
    rho = torch.tensor([0.4], dtype=torch.float64)
    xbar = 2.0 * (y - g)
    full_x, full = O.layer_grads(y, xbar, h, None, lam, rho, K, False)
    assert rel_l2(out["weight"].permute(3, 2, 1, 0).double(), full["weight"]) < 1e-5
    assert abs(float(out["lam"]) - float(full["lam"])) < 1e-4 * abs(float(full["lam"]))
    assert abs(float(out["rho"]) - float(full["rho"])) < 1e-4 * abs(float(full["rho"]))
    assert rel_l2(out["x0"], full_x[..., 0:2]) < 1e-12      # forward of a shard == that slice of the full batch


# ---- exact global-batch isotropic TV across ranks: emulated kernels + gloo ---------------------
def _iso_worker(rank, world, port, out, iso_flag):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu_harness as E
        from admm_deconv_b200 import _lib
        torch.set_num_threads(1)
        lib = E.emu_lib()
        M, N, P, B, K = 32, 32, 2, 3, 3
        y, h, g = make_case(M, N, P, B, 3, 3, 77)
        xbar = torch.from_numpy(np.random.default_rng(5).standard_normal((M, N, P, B)))
        lo, hi = D.shard_range(B, rank, world)
        ys = E.f32(y[..., lo:hi].numpy()); xb = E.f32(xbar[..., lo:hi].numpy())
        Bs = hi - lo
        d = _lib.make_desc(M, N, P, Bs, 3, 3, K, True, "identity", False, 0, _lib.FLAG_NO_CLAMP | iso_flag, 0.0)
        fwd_b, ck_b, bwd_b = lib.workspace_bytes(d)
        ws, ck, wsb = (torch.from_numpy(E.aligned_bytes(n)) for n in (fwd_b, ck_b, bwd_b))
        hb = E.f32(h.numpy()[:, :, 0, 0]); lam = np.array([0.05], np.float32); rho = np.array([0.3], np.float32)
        x = np.asfortranarray(np.zeros((M, N, P, Bs), np.float32))
        cp = D.IsoCoupling()
        cp.register(ws, ck)
        lib.forward_ex(d, E.ptr(ys), E.ptr(hb), E.ptr(lam), E.ptr(rho), None, E.ptr(x), ws.data_ptr(), ck.data_ptr(), None, cp.hooks)
        assert cp.calls == K - 1 and cp.floats == (K - 1) * M * N
        ybar = np.asfortranarray(np.zeros((M, N, P, Bs), np.float32))
        hbar = np.asfortranarray(np.zeros((3, 3), np.float32)); lb = np.zeros(1, np.float32); rb = np.zeros(1, np.float32)
        cp.register(wsb, ck)
        lib.backward_ex(d, E.ptr(xb), E.ptr(x), E.ptr(ys), E.ptr(hb), E.ptr(lam), E.ptr(rho), ck.data_ptr(), E.ptr(ybar),
                        E.ptr(hbar), E.ptr(lb), E.ptr(rb), None, wsb.data_ptr(), None, cp.hooks)
        packed = torch.from_numpy(np.concatenate([hbar.reshape(-1, order="F"), lb, rb]).astype(np.float32))
        dist.all_reduce(packed)                       # the usual one-call gradient all-reduce
        out[f"x{rank}"] = torch.from_numpy(np.array(x)); out[f"ybar{rank}"] = torch.from_numpy(np.array(ybar))
        if rank == 0:
            out["packed"] = packed.clone()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("iso_flag", [16, 32])    # precomputed / inline per-pixel terms
def test_two_rank_global_isotropic_equals_single_device(iso_flag, emu):
    # `emu`: the emulation library is built HERE, once, before the two ranks start (they would otherwise race to build it)
    world = 2
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_iso_worker, args=(world, port, out, iso_flag), nprocs=world, join=True)
    M, N, P, B, K = 32, 32, 2, 3, 3
    y, h, g = make_case(M, N, P, B, 3, 3, 77)
    xbar = torch.from_numpy(np.random.default_rng(5).standard_normal((M, N, P, B)))
    y32, h32 = y.float().double(), h.float().double()
    lam = torch.tensor([0.05], dtype=torch.float32).double(); rho = torch.tensor([0.3], dtype=torch.float32).double()
    full_x, full = O.layer_grads(y32, xbar.float().double(), h32, None, lam, rho, K, True)
    x = torch.cat([out["x0"], out["x1"]], dim=3).double()
    ybar = torch.cat([out["ybar0"], out["ybar1"]], dim=3).double()
    assert rel_l2(x, full_x) < 1e-5                                  # the whole-batch reference result
    # per-shard isotropic norms give a different x: the coupling is what makes it match
    x_shard = torch.cat([O.tvd_fft_cpu(y32[..., :2], lam, rho, h32, True, K), O.tvd_fft_cpu(y32[..., 2:], lam, rho, h32, True, K)], dim=3)
    assert rel_l2(x_shard, full_x) > 1e-4
    assert rel_l2(ybar, full["x"]) < 1e-4
    pk = out["packed"].double()
    assert rel_l2(pk[:9].reshape(3, 3).t().reshape(3, 3, 1, 1), full["weight"]) < 1e-4
    assert abs(float(pk[9]) - float(full["lam"])) < 2e-4 * abs(float(full["lam"]))
    assert abs(float(pk[10]) - float(full["rho"])) < 2e-4 * abs(float(full["rho"]))
