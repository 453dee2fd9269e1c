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
