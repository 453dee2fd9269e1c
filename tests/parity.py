"""Shared parity checks (backend-agnostic: see harness.py)."""
from __future__ import annotations

import numpy as np
import torch

from cases import rel_l2
from oracle import admm_tv_oracle as O
from oracle import teacher_forced as TF

T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()


def act_grad(out: torch.Tensor, act: str) -> torch.Tensor:
    if act == "relu":
        return (out > 0).double()
    if act == "relu6":
        return ((out > 0) & (out < 6)).double()
    if act == "relu1":
        return ((out > 0) & (out < 1)).double()
    return torch.ones_like(out)


def close(a: float, b: float, rtol: float, floor: float = 1e-3) -> bool:
    return abs(a - b) <= rtol * max(abs(b), floor)


def check_forward(be, y, h, lam, rho, iso, K, tol=1e-5, fast=True, **kw):
    """Device forward vs the fp64 oracle on fp32-rounded inputs.  Returns the rel-L2 error."""
    y = y.float().double()
    h = None if h is None else h.float().double()
    r = be.forward(y.numpy(), lam, rho, None if h is None else h.numpy()[:, :, 0, 0], iso, K, flags=1, **kw)
    lt = torch.tensor(np.atleast_1d(lam), dtype=torch.float32).double()
    rt = torch.tensor(np.atleast_1d(rho), dtype=torch.float32).double()
    xo = (O.tvd_fft_fast if fast or lt.numel() > 1 else O.tvd_fft_cpu)(y, lt, rt, h, iso, K)
    err = rel_l2(T(r["x"].get()), xo)
    assert err < tol, err
    return err


def check_backward(be, y, h, lam, rho, iso, K, xbar, act="identity", bias=None, creg=0.0, flags=0,
                   tol=1e-5, tol_scalar=1e-4, tol_e2e=None):
    """Device forward+backward vs (1) the teacher-forced fp64 adjoint replaying the device's own
    checkpointed states (arithmetic parity, tol) and optionally (2) end-to-end fp64 autograd (tol_e2e).
    Returns a dict of the measured errors and the number of threshold-mask flips."""
    y = y.float().double()
    h = None if h is None else h.float().double()
    xbar = xbar.float().double()
    f = be.forward(y.numpy(), lam, rho, None if h is None else h.numpy()[:, :, 0, 0], iso, K, act=act, bias=bias, creg=creg,
                   flags=flags, want_ckpt=True)
    g = be.backward(f, xbar.numpy())
    # parameters as the device left them (clamped, fp32)
    lt = T(f["lam"].get()); rt = T(f["rho"].get())
    hc = None if h is None else T(f["h"].get()).reshape(h.shape)
    x_dev = T(f["x"].get())
    states = be.ckpt_states(f)
    # isotropic: the teacher-forced adjoint also replays the device's per-pixel norms, hence its gates n > tau
    norms = be.ckpt_norms(f) if (iso and K > 1) else None
    xbar_eff = xbar * act_grad(x_dev, act)
    tf = TF.backward(xbar_eff, y, lt, rt, hc, iso, K, states, nograd_repeat=bool(flags & 2), nsq_states=norms, fp32_gate=True)
    # the same recursion evaluated in the device's working precision (fp32 arrays, complex64 FFTs, fp32 sums): its distance
    # from the fp64 result is the rounding floor of the cancelling scalar sums (lambdabar, rhobar) and of hbar
    tf32 = TF.backward(xbar_eff, y, lt, rt, hc, iso, K, states, nograd_repeat=bool(flags & 2), nsq_states=norms, fp32_gate=True,
                       dtype=torch.float32)
    relf = lambda a, b: float((a.double().reshape(-1) - b.reshape(-1)).abs().max() / max(float(b.abs().max()), 1e-3))
    res = {"gate_margin": tf["gate_margin"], "lam_floor32": relf(tf32["lam"], tf["lam"]), "rho_floor32": relf(tf32["rho"], tf["rho"])}
    if h is not None:
        res["hbar_floor32"] = rel_l2(tf32["weight"].double(), tf["weight"])
    res["ybar"] = rel_l2(T(g["ybar"]), tf["x"])
    assert res["ybar"] < tol, ("ybar", res["ybar"])
    # scalars: the worst entry (one per iteration with ADMMTV_FLAG_PER_ITER_PARAMS), each relative to the largest entry
    sc_err = lambda a, b: float((T(a).reshape(-1) - b.reshape(-1)).abs().max() / max(float(b.abs().max()), 1e-3))
    if K > 1:
        res["lam"] = sc_err(g["lambar"], tf["lam"])
        assert res["lam"] < tol_scalar, ("lambar", g["lambar"], tf["lam"])
    res["rho"] = sc_err(g["rhobar"], tf["rho"])
    assert res["rho"] < tol_scalar, ("rhobar", g["rhobar"], tf["rho"])
    if h is not None:
        # clamp gate (deconv_admm.jl:219): gradient only where 0 <= h <= 1 before the clamp
        gate = ((h >= 0) & (h <= 1)).double() if not (flags & 1) else torch.ones_like(h)
        res["hbar"] = rel_l2(T(g["hbar"]).reshape(h.shape), tf["weight"] * gate)
        assert res["hbar"] < tol, ("hbar", res["hbar"])
    if bias is not None:
        assert close(float(g["biasbar"][0]), float(xbar_eff.sum()), 1e-5)
    # mask flips of the device forward relative to the fp64 forward, and end-to-end agreement
    _, st64 = TF.forward_states(y, lt, rt, hc, iso, K)
    if lt.numel() == 1:
        if not iso:
            res["flips"] = TF.count_mask_flips(states, st64, float(lt / rt))
        elif K > 1:
            res["flips"] = TF.count_gate_flips_iso(norms, st64, float(lt / rt))
    if tol_e2e is not None:
        bt = None if bias is None else torch.tensor([bias], dtype=torch.float32).double()
        _, go = O.layer_grads(y, xbar, None if h is None else T(f["h"].get()).reshape(h.shape), bt, lt, rt, K, iso, 0.0, act,
                              nograd_repeat=bool(flags & 2))
        res["ybar_e2e"] = rel_l2(T(g["ybar"]), go["x"])
        assert res["ybar_e2e"] < tol_e2e, ("ybar end-to-end", res["ybar_e2e"], res.get("flips"))
    return res
