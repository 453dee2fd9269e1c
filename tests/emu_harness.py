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

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libadmmtv_emu.so")
CSRC = os.path.join(ROOT, "admm_deconv_b200", "csrc")


def _sources():
    out = [os.path.join(EMU_DIR, f) for f in ("cuda_emu.h", "cuda_emu.cpp")]
    out += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".inc", ".h"))]
    out.append(os.path.join(ROOT, "include", "admmtv.h"))
    return out


def build_emu(force: bool = False) -> str:
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    if not force and os.path.exists(EMU_SO):
        t = os.path.getmtime(EMU_SO)
        if all(os.path.getmtime(s) <= t for s in _sources()):
            return EMU_SO
    cus = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    cmd = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-DADMMTV_EMU", "-I" + EMU_DIR, "-I" + CSRC,
           "-x", "c++", *cus, os.path.join(EMU_DIR, "cuda_emu.cpp"), "-o", EMU_SO]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return EMU_SO


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
