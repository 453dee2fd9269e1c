"""TEST INFRASTRUCTURE: builds and drives the CPU *emulation* build of the kernel sources
(tests/emu/cuda_emu.h) through the same C ABI, with numpy buffers standing in for device memory.
Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from admm_deconv_b200 import _lib  # noqa: E402

from admm_deconv_b200 import build as _build  # noqa: E402


def build_emu(force: bool = False) -> str:
    return _build.build(emulate=True, force=force)


_EMU = None


def emu_lib() -> _lib.AdmmTvLib:
    global _EMU
    if _EMU is None:
        _EMU = _lib.AdmmTvLib(build_emu())
    return _EMU


def f32(a) -> np.ndarray:
    """Julia (column-major) fp32 buffer of an (M,N,P,B)- or (kh,kw)-indexed array."""
    return np.asfortranarray(np.asarray(a, dtype=np.float32))


def ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def aligned_bytes(n: int) -> np.ndarray:
    raw = np.zeros(n + 256, dtype=np.uint8)
    off = (-raw.ctypes.data) % 256
    return raw[off:off + n]


def forward(lib, y, lam, rho, h=None, iso=False, iters=10, act="identity", bias=None, creg=0.0, flags=0,
            want_ckpt=False):
    """Runs admmtv_forward on numpy buffers.  Returns dict(x, lam, rho, h, ckpt, desc, ws...)."""
    y = f32(y)
    M, N, P, B = y.shape
    kh, kw = (0, 0) if h is None else (h.shape[0], h.shape[1])
    d = _lib.make_desc(M, N, P, B, kh, kw, iters, iso, act, bias is not None, 0, flags, creg)
    fwd_b, ck_b, bwd_b = lib.workspace_bytes(d)
    ws = aligned_bytes(fwd_b)
    ck = aligned_bytes(ck_b) if want_ckpt else None
    hbuf = None if h is None else f32(np.asarray(h).reshape(kh, kw))
    lbuf = np.array([lam], dtype=np.float32).reshape(1)
    rbuf = np.array([rho], dtype=np.float32).reshape(1)
    bbuf = None if bias is None else np.array([bias], dtype=np.float32).reshape(1)
    x = np.zeros((M, N, P, B), dtype=np.float32, order="F")
    lib.forward(d, ptr(y), None if hbuf is None else ptr(hbuf), ptr(lbuf), ptr(rbuf),
                None if bbuf is None else ptr(bbuf), ptr(x), ptr(ws), None if ck is None else ptr(ck), None)
    return dict(x=x, lam=lbuf, rho=rbuf, h=hbuf, ckpt=ck, desc=d, y=y, bias=bbuf, bwd_bytes=bwd_b)


def backward(lib, fwd, xbar):
    """Runs admmtv_backward on the buffers a `forward(..., want_ckpt=True)` call returned."""
    d = fwd["desc"]
    M, N, P, B = d.M, d.N, d.P, d.B
    xbar = f32(xbar)
    ws = aligned_bytes(fwd["bwd_bytes"])
    ybar = np.zeros((M, N, P, B), dtype=np.float32, order="F")
    hbar = None if fwd["h"] is None else np.zeros_like(fwd["h"])
    lbar = np.zeros(1, dtype=np.float32)
    rbar = np.zeros(1, dtype=np.float32)
    bbar = None if fwd["bias"] is None else np.zeros(1, dtype=np.float32)
    lib.backward(d, ptr(xbar), ptr(fwd["x"]), ptr(fwd["y"]), None if fwd["h"] is None else ptr(fwd["h"]), ptr(fwd["lam"]),
                 ptr(fwd["rho"]), ptr(fwd["ckpt"]), ptr(ybar), None if hbar is None else ptr(hbar), ptr(lbar), ptr(rbar),
                 None if bbar is None else ptr(bbar), ptr(ws), None)
    return dict(ybar=ybar, hbar=hbar, lambar=lbar, rhobar=rbar, biasbar=bbar)
