"""Teacher-forced fp64 backward  --  TEST INFRASTRUCTURE ONLY (see admm_tv_oracle.py header).

End-to-end gradient parity between an fp32 and an fp64 run is limited by a handful of pixels whose
|v| sits within rounding of the threshold tau and flips the shrinkage mask 1[|v|>tau]
 -- not by arithmetic.  To test the backward ARITHMETIC at the 1e-5
level this module evaluates the exact adjoint recursion of SURVEY.md 8a-10 in fp64 while replaying
the per-iteration states v_k = D x_k + u_{k-1} that the device forward saved (its checkpoint), so
both sides use identical masks.  The recursion itself is verified against torch.autograd through
the literal restatement (tests/test_oracle.py::test_teacher_forced_equals_autograd...).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

from . import admm_tv_oracle as O

DT = torch.float64


def _full_tables(M, N, h, rho, dt=DT):
    k1 = torch.arange(M, dtype=DT).reshape(-1, 1)
    k2 = torch.arange(N, dtype=DT).reshape(1, -1)
    L = (4 * torch.sin(math.pi * k2 / N) ** 2 + 4 * torch.sin(math.pi * k1 / M) ** 2).to(dt)
    if h is None or h.numel() == 0:
        Sig = torch.ones(M, N, dtype=torch.complex128 if dt == DT else torch.complex64)
    else:
        hh = torch.zeros(M, N, dtype=dt)
        hh[: h.shape[0], : h.shape[1]] = h[:, :, 0, 0]
        Sig = torch.fft.fftn(hh)
    C = 1.0 / (Sig.abs() ** 2 + rho * L)
    return Sig, L, C


def forward_states(y, lam, rho, h, iso, K):
    """fp64 forward in v-state form; returns (x_K, [v_1..v_{K-1}] as (v1, v2) pairs)."""
    st: list = []
    x = O.tvd_fft_fast(y, lam, rho, h, iso, K, states=st)
    return x, [(s[1], s[2]) for s in st[:-1]]


def backward(xbar, y, lam, rho, h, iso, K, v_states: List[Tuple[torch.Tensor, torch.Tensor]], nograd_repeat=False,
             nsq_states: Optional[List[torch.Tensor]] = None, fp32_gate: bool = False, dtype=DT):
    """Exact adjoint given the states v_1..v_{K-1} (each (M,N,P,B) fp64 pair).  Returns
    dict(x=ybar, lam, rho, weight, gate_margin).

    Teacher forcing covers every DECISION of the replayed forward, not only its states:
      * nsq_states -- isotropic: the per-pixel |v_k|^2 (M,N) that the device forward accumulated in fp32 and
        checkpointed (k = 1..K-1).  The norm n = sqrt_fp32(nsq) and the gate n > tau are then the device's own; without
        it the norm is recomputed in fp64 from v, and a pixel with n within rounding of tau can take the other branch
        of BT (ops.jl:10), whose derivative coefficient tau <v,q> / n^3 is discontinuous there.
      * fp32_gate -- the comparisons |v| > tau (ST) and n > tau (BT) use tau rounded as the device rounds it
        (fp32 lambda / rho); the arithmetic keeps the fp64 tau.
    gate_margin = the smallest relative distance of a decision variable from tau (how close a flip was).

    dtype = torch.float32 evaluates the SAME recursion with fp32 arrays and complex64 FFTs (reductions included): its
    distance from the fp64 result is the rounding floor of this computation in the device's working precision, which is
    what the scalar gradients (cancelling sums) are compared against in tests/parity.py."""
    M, N, P, B = y.shape
    dt = dtype
    PS = lam.numel()                  # 1, or K per-iteration values (EXTENSION: entry k-1 belongs to iteration k)
    pe = lambda k: 0 if PS == 1 else k - 1
    tau_g64_all = (lam.float() / rho.float()).double() if fp32_gate else (lam / rho).double()
    xbar, y, lam, rho = xbar.to(dt), y.to(dt), lam.to(dt), rho.to(dt)
    h = None if h is None else h.to(dt)
    v_states = [(a.to(dt), b.to(dt)) for a, b in v_states]
    tau_all = lam / rho
    margin = float("inf")
    tabs = [_full_tables(M, N, h, rho[e:e + 1], dt) for e in range(PS)]
    Sig, L = tabs[0][0], tabs[0][1]
    fft2 = lambda t: torch.fft.fftn(t, dim=(0, 1))
    ifft2 = lambda T: torch.fft.ifftn(T, dim=(0, 1)).real
    b = y if h is None or h.numel() == 0 else O.Ht_roll(y, h)

    def shrink(v1, v2, k):
        tau, tau_g = tau_all[pe(max(k, 1))], tau_g64_all[pe(max(k, 1))].to(dt)   # v_k was shrunk with tau_k
        if iso:
            if nsq_states is not None and k >= 1:
                n = torch.sqrt(nsq_states[k - 1].float()).to(dt).reshape(M, N, 1, 1)   # the device's fp32 norm
            else:
                n = torch.sqrt(torch.sum(v1 * v1 + v2 * v2, dim=(2, 3), keepdim=True))
            s = torch.where(n > 0, torch.clamp(1 - tau / n, min=0), torch.zeros_like(n))
            s = torch.where(n > tau_g, s, torch.zeros_like(s))   # the gate as the device takes it
            return s * v1, s * v2, n, s
        m1, m2 = v1.abs() > tau_g, v2.abs() > tau_g
        return torch.where(m1, v1 - torch.sign(v1) * tau, torch.zeros_like(v1)), torch.where(m2, v2 - torch.sign(v2) * tau, torch.zeros_like(v2)), None, None

    zero = torch.zeros_like(y)
    vs = [(zero, zero)] + list(v_states)          # vs[k] = v_k, v_0 = 0
    vb1, vb2 = zero, zero
    G = torch.zeros(PS, M, N, dtype=dt)
    bbar = torch.zeros_like(y)
    rhobar = torch.zeros(PS, dtype=dt)
    taubar = torch.zeros(PS, dtype=dt)
    for k in range(K, 0, -1):
        xk = (xbar if k == K else 0) + (O.Dt_roll(vb1, vb2) if k < K else 0)
        v1, v2 = vs[k - 1]
        z1, z2, n, s = shrink(v1, v2, k - 1)
        g1, g2 = 2 * z1 - v1, 2 * z2 - v2
        rho_k = rho[pe(k)]
        r = b + rho_k * O.Dt_roll(g1, g2)
        Zb = fft2(xk)
        G[pe(k)] += (Zb.conj() * fft2(r)).real.sum(dim=(2, 3))
        rb = ifft2(tabs[pe(k)][2].reshape(M, N, 1, 1) * Zb)
        bbar = bbar + rb
        if k == 1:
            break
        d1, d2 = O.D_roll(rb)
        rhobar[pe(k)] = rhobar[pe(k)] + (d1 * g1).sum() + (d2 * g2).sum()
        gb1, gb2 = rho_k * d1, rho_k * d2
        tau, tau_g = tau_all[pe(k - 1)], tau_g64_all[pe(k - 1)].to(dt)     # the shrinkage of v_{k-1}
        te = pe(k - 1)
        q1, q2 = 2 * gb1 - vb1, 2 * gb2 - vb2
        if iso:
            ip = torch.sum(q1 * v1 + q2 * v2, dim=(2, 3), keepdim=True)
            act = n > tau_g
            margin = min(margin, float(((n - tau_g).abs() / tau_g).min()))
            coef = torch.where(act, tau * ip / n ** 3, torch.zeros_like(n))
            nv1 = vb1 - gb1 + s * q1 + coef * v1
            nv2 = vb2 - gb2 + s * q2 + coef * v2
            taubar[te] = taubar[te] - torch.where(act, ip / n, torch.zeros_like(n)).sum()
        else:
            m1 = (v1.abs() > tau_g).to(dt)
            m2 = (v2.abs() > tau_g).to(dt)
            margin = min(margin, float(((v1.abs() - tau_g).abs() / tau_g).min()), float(((v2.abs() - tau_g).abs() / tau_g).min()))
            nv1 = vb1 - gb1 + m1 * q1
            nv2 = vb2 - gb2 + m2 * q2
            taubar[te] = taubar[te] - (torch.sign(v1) * m1 * q1).sum() - (torch.sign(v2) * m2 * q2).sum()
        vb1, vb2 = nv1, nv2
    Sbar_e = torch.stack([-(G[e] / (M * N)) * tabs[e][2] * tabs[e][2] for e in range(PS)])
    rhobar = rhobar + (Sbar_e * L).sum(dim=(1, 2))
    Sbar = Sbar_e.sum(dim=0)
    out = {"lam": (taubar / rho).reshape(PS), "rho": (rhobar - taubar * lam / rho ** 2).reshape(PS), "gate_margin": margin}
    if h is None or h.numel() == 0:
        out["x"] = bbar
        out["weight"] = None
        return out
    kh, kw = h.shape[:2]
    pd, pr = (kh - 1) // 2, (kw - 1) // 2
    out["x"] = O.H_forward(bbar, h)
    Fh = torch.fft.ifftn(2 * Sbar * Sig) * (M * N)
    hb = torch.zeros(kh, kw, dtype=dt)
    for a in range(kh):
        for c in range(kw):
            hb[a, c] = Fh[a, c].real
            if not nograd_repeat:
                hb[a, c] += (bbar * torch.roll(y, shifts=(-(a - pd), -(c - pr)), dims=(0, 1))).sum()
    out["weight"] = hb.reshape(kh, kw, 1, 1)
    return out


def count_gate_flips_iso(nsq_dev: List[torch.Tensor], v_ref, tau: float) -> int:
    """Isotropic: how many per-pixel decisions n > tau differ between the device's checkpointed norms and the
    norms of a reference state list (one decision per pixel and iteration)."""
    n = 0
    for nd, (b1, b2) in zip(nsq_dev, v_ref):
        nr = torch.sqrt(torch.sum(b1 * b1 + b2 * b2, dim=(2, 3)))
        n += int(((torch.sqrt(nd.double()) > tau) != (nr > tau)).sum())
    return n


def count_mask_flips(v_a, v_b, tau: float) -> int:
    """How many shrinkage decisions differ between two state lists (aniso)."""
    n = 0
    for (a1, a2), (b1, b2) in zip(v_a, v_b):
        n += int(((a1.abs() > tau) != (b1.abs() > tau)).sum()) + int(((a2.abs() > tau) != (b2.abs() > tau)).sum())
    return n
