"""Generates the golden fixtures in this directory from the fp64 oracle (oracle/admm_tv_oracle.py).

The reference cannot run here (no Julia; its only test has no stored values -- SURVEY.md 8c), so
these vectors pin the *oracle's* outputs, not the reference's: PARITY UNPINNED.  They exist so a
later change to the oracle or the kernels is caught, and so the GPU box (which has no
/root/reference and need not re-run the slow fp64 oracle for these) compares against fixed files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import make_case  # noqa: E402
from oracle import admm_tv_oracle as O  # noqa: E402

# name, M, N, P, B, kh, kw, psf, iso, iters, lam, rho, act, bias, creg
CASES = [
    ("aniso_64x64_rgb2_k7line", 64, 64, 3, 2, 7, 7, "line", False, 20, 0.0041, 0.021, "identity", None, 0.0),
    ("aniso_32x64_g3_nopsf", 32, 64, 1, 3, 0, 0, None, False, 12, 0.05, 0.3, "relu1", 0.01, 0.0),
    ("aniso_64x32_g2_k4x7", 64, 32, 1, 2, 4, 7, "random", False, 10, 0.02, 0.1, "relu6", None, 0.0),
    ("aniso_128x128_g1_gauss9", 128, 128, 1, 1, 9, 9, "gauss", False, 30, 0.0041, 0.021, "identity", None, 0.0),
    ("iso_64x64_rgb2_k5", 64, 64, 3, 2, 5, 5, "random", True, 15, 0.02, 0.1, "identity", None, 0.0),
    ("iso_32x32_g4_nopsf", 32, 32, 1, 4, 0, 0, None, True, 10, 0.05, 0.3, "relu", None, 0.0),
]


def main():
    for i, (name, M, N, P, B, kh, kw, psf, iso, K, lam, rho, act, bias, creg) in enumerate(CASES):
        y, h, g = make_case(M, N, P, B, kh, kw, 1000 + i, psf or "random")
        y32 = y.float().double()            # the GPU sees fp32 inputs: round first, then run the fp64 oracle
        h32 = None if h is None else h.float().double()
        lam_t = torch.tensor([lam], dtype=torch.float32).double()
        rho_t = torch.tensor([rho], dtype=torch.float32).double()
        b_t = None if bias is None else torch.tensor([bias], dtype=torch.float32).double()
        xbar = torch.from_numpy(np.random.Generator(np.random.PCG64(77 + i)).standard_normal((M, N, P, B))).float().double()
        out, grads = O.layer_grads(y32, xbar, h32, b_t, lam_t, rho_t, K, iso, creg, act)
        d = dict(y=y32.numpy().astype(np.float32), x=out.numpy(), xbar=xbar.numpy().astype(np.float32),
                 lam=np.float32(lam), rho=np.float32(rho), iso=iso, iters=K, act=act, creg=creg,
                 ybar=grads["x"].numpy(), lambar=grads["lam"].numpy(), rhobar=grads["rho"].numpy())
        if h is not None:
            d["h"] = h32.numpy().astype(np.float32)
            d["hbar"] = grads["weight"].numpy()
        if bias is not None:
            d["bias"] = np.float32(bias)
            d["biasbar"] = grads["bias"].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "x range", float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
