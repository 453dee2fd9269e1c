"""TEST INFRASTRUCTURE: one driver for the C ABI with two backends --
  EmuBackend : the g++ emulation build of the kernel sources, numpy buffers (CPU suite)
  GpuBackend : the nvcc-built product library, torch CUDA buffers (-m gpu suite)
Both return numpy arrays in the Julia (M,N,P,B) index order."""
from __future__ import annotations

import numpy as np
import torch

from admm_deconv_b200 import _lib


def f32(a) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=np.float32))


class _Buf:
    """A typed buffer living on the backend's 'device'."""

    def __init__(self, be, arr: np.ndarray):
        self.be, self.shape, self.dtype = be, arr.shape, arr.dtype
        flat = np.ascontiguousarray(arr.reshape(-1, order="F"))
        if be.gpu:
            self.t = torch.from_numpy(flat.view(np.uint8).copy()).cuda()
            self.ptr = self.t.data_ptr()
        else:
            raw = np.zeros(flat.nbytes + 256, dtype=np.uint8)
            off = (-raw.ctypes.data) % 256
            self.t = raw[off:off + flat.nbytes]
            self.t[:] = flat.view(np.uint8)
            self.ptr = self.t.ctypes.data

    def get(self) -> np.ndarray:
        host = self.t.cpu().numpy() if self.be.gpu else self.t
        return host.view(self.dtype).reshape(self.shape, order="F").copy()


class Backend:
    gpu = False

    def __init__(self, lib):
        self.lib = lib

    def buf(self, arr):
        return _Buf(self, np.asarray(arr))

    def zeros(self, shape, dtype=np.float32):
        return _Buf(self, np.zeros(shape, dtype=dtype))

    def sync(self):
        if self.gpu:
            torch.cuda.synchronize()

    def stream(self):
        return torch.cuda.current_stream().cuda_stream if self.gpu else None

    # ------------------------------------------------------------------------------------
    def forward(self, y, lam, rho, h=None, iso=False, iters=10, act="identity", bias=None, creg=0.0, flags=0,
                want_ckpt=False):
        y = f32(y)
        M, N, P, B = y.shape
        kh, kw = (0, 0) if h is None else (h.shape[0], h.shape[1])
        lam = np.atleast_1d(np.asarray(lam, dtype=np.float32)); rho = np.atleast_1d(np.asarray(rho, dtype=np.float32))
        if lam.size > 1 or rho.size > 1:      # one (lambda, rho) per iteration: ADMMTV_FLAG_PER_ITER_PARAMS
            assert lam.size == iters and rho.size == iters
            flags |= _lib.FLAG_PER_ITER_PARAMS
        d = _lib.make_desc(M, N, P, B, kh, kw, iters, iso, act, bias is not None, 0, flags, creg)
        fwd_b, ck_b, bwd_b = self.lib.workspace_bytes(d)
        r = dict(desc=d, bwd_bytes=bwd_b)
        r["y"] = self.buf(y)
        r["h"] = None if h is None else self.buf(f32(np.asarray(h).reshape(kh, kw)))
        r["lam"] = self.buf(lam)
        r["rho"] = self.buf(rho)
        r["bias"] = None if bias is None else self.buf(np.array([bias], dtype=np.float32))
        r["x"] = self.zeros((M, N, P, B))
        ws = self.zeros((fwd_b,), np.uint8)
        r["ckpt"] = self.zeros((ck_b,), np.uint8) if want_ckpt else None
        p = lambda b: None if b is None else b.ptr
        self.lib.forward(d, p(r["y"]), p(r["h"]), p(r["lam"]), p(r["rho"]), p(r["bias"]), p(r["x"]), ws.ptr, p(r["ckpt"]),
                         self.stream())
        self.sync()
        return r

    def forward_grouped(self, y, lam, rho, h=None, iso=False, iters=10, groups=1, shared_input=False, concat=False,
                        act="identity", want_ckpt=False):
        """y: (M,N,P,Bin) Julia-indexed; lam, rho: length-G; h: (kh,kw,G) or None.  Returns x, or with
        want_ckpt the dict `backward_grouped` takes."""
        y = f32(y)
        M, N, P, Bin = y.shape
        Bg = Bin if shared_input else Bin // groups
        kh, kw = (0, 0) if h is None else (h.shape[0], h.shape[1])
        flags = _lib.FLAG_NO_CLAMP | (_lib.FLAG_SHARED_INPUT if shared_input else 0) | (_lib.FLAG_CHANNEL_CONCAT if concat else 0)
        d = _lib.make_desc(M, N, P, groups * Bg, kh, kw, iters, iso, act, False, 0, flags, 0.0, groups)
        fwd_b, ck_b, bwd_b = self.lib.workspace_bytes(d)
        yb = self.buf(y)
        hb = None if h is None else self.buf(f32(h))
        lb = self.buf(np.asarray(lam, dtype=np.float32))
        rb = self.buf(np.asarray(rho, dtype=np.float32))
        shape = (M, N, groups * P, Bg) if concat else (M, N, P, groups * Bg)
        x = self.zeros(shape)
        ws = self.zeros((fwd_b,), np.uint8)
        ck = self.zeros((ck_b,), np.uint8) if want_ckpt else None
        self.lib.forward(d, yb.ptr, None if hb is None else hb.ptr, lb.ptr, rb.ptr, None, x.ptr, ws.ptr,
                         None if ck is None else ck.ptr, self.stream())
        self.sync()
        if not want_ckpt:
            return x.get()
        return dict(desc=d, y=yb, h=hb, lam=lb, rho=rb, x=x, ckpt=ck, bwd_bytes=bwd_b, groups=groups, in_shape=y.shape)

    def backward_grouped(self, fwd, xbar):
        d, G = fwd["desc"], fwd["groups"]
        xb = self.buf(f32(xbar))
        ws = self.zeros((fwd["bwd_bytes"],), np.uint8)
        ybar = self.zeros(fwd["in_shape"])
        hbar = None if fwd["h"] is None else self.zeros(fwd["h"].shape)
        lbar, rbar = self.zeros((G,)), self.zeros((G,))
        p = lambda b: None if b is None else b.ptr
        self.lib.backward(d, xb.ptr, fwd["x"].ptr, fwd["y"].ptr, p(fwd["h"]), fwd["lam"].ptr, fwd["rho"].ptr, fwd["ckpt"].ptr,
                          ybar.ptr, p(hbar), lbar.ptr, rbar.ptr, None, ws.ptr, self.stream())
        self.sync()
        return dict(ybar=ybar.get(), hbar=None if hbar is None else hbar.get(), lambar=lbar.get(), rhobar=rbar.get())

    def backward(self, fwd, xbar):
        d = fwd["desc"]
        M, N, P, B = d.M, d.N, d.P, d.B
        xb = self.buf(f32(xbar))
        ws = self.zeros((fwd["bwd_bytes"],), np.uint8)
        out = dict(ybar=self.zeros((M, N, P, B)), hbar=None if fwd["h"] is None else self.zeros(fwd["h"].shape),
                   lambar=self.zeros(fwd["lam"].shape), rhobar=self.zeros(fwd["rho"].shape),
                   biasbar=None if fwd["bias"] is None else self.zeros((1,)))
        p = lambda b: None if b is None else b.ptr
        self.lib.backward(d, xb.ptr, fwd["x"].ptr, fwd["y"].ptr, p(fwd["h"]), fwd["lam"].ptr, fwd["rho"].ptr, fwd["ckpt"].ptr,
                          out["ybar"].ptr, p(out["hbar"]), out["lambar"].ptr, out["rhobar"].ptr, p(out["biasbar"]), ws.ptr,
                          self.stream())
        self.sync()
        return {k: (None if v is None else v.get()) for k, v in out.items()}

    def backward_mse(self, fwd, target):
        """admmtv_backward_mse: the MSE cotangent 2 (x - target) / numel is formed inside the first kernel; also returns the loss."""
        d = fwd["desc"]
        M, N, P, B = d.M, d.N, d.P, d.B
        tb = self.buf(f32(target))
        ws = self.zeros((fwd["bwd_bytes"],), np.uint8)
        loss = self.zeros((1,), np.float64)
        out = dict(ybar=self.zeros((M, N, P, B)), hbar=None if fwd["h"] is None else self.zeros(fwd["h"].shape),
                   lambar=self.zeros(fwd["lam"].shape), rhobar=self.zeros(fwd["rho"].shape),
                   biasbar=None if fwd["bias"] is None else self.zeros((1,)))
        p = lambda b: None if b is None else b.ptr
        self.lib.backward_mse(d, tb.ptr, fwd["x"].ptr, fwd["y"].ptr, p(fwd["h"]), fwd["lam"].ptr, fwd["rho"].ptr, fwd["ckpt"].ptr,
                              out["ybar"].ptr, p(out["hbar"]), out["lambar"].ptr, out["rhobar"].ptr, p(out["biasbar"]), loss.ptr, ws.ptr,
                              self.stream())
        self.sync()
        res = {k: (None if v is None else v.get()) for k, v in out.items()}
        res["loss"] = float(loss.get()[0]) / (M * N * P * B)
        return res

    def ckpt_states(self, fwd):
        """[(v1_k, v2_k)] for k = 1..K-1 as fp64 torch (M,N,P,B) arrays, read from the checkpoint."""
        d = fwd["desc"]
        M, N, P, B, K = d.M, d.N, d.P, d.B, d.iters
        S = P * B
        Q = (S + 1) // 2
        off = self.lib.ckpt_layout(d)
        raw = fwd["ckpt"].get()
        n = (K - 1) * Q * 2 * N * M * 2
        v = raw[off[1]:off[1] + 4 * n].view(np.float32).reshape(K - 1, Q, 2, N, M, 2)
        states = []
        for k in range(K - 1):
            chans = []
            for ch in range(2):
                a = v[k, :, ch]                      # (Q, N, M, 2)
                planes = np.transpose(a, (2, 1, 0, 3)).reshape(M, N, 2 * Q)[:, :, :S]   # plane index = 2q + c
                chans.append(torch.from_numpy(planes.reshape(M, N, P, B, order="F").astype(np.float64)))
            states.append((chans[0], chans[1]))
        return states


    def ckpt_norms(self, fwd):
        """Isotropic: the per-pixel |v_k|^2 the device accumulated (fp32, fixed-order sum over the plane pairs) and
        checkpointed, k = 1..K-1, as (M,N) fp32 torch arrays (single-group calls)."""
        d = fwd["desc"]
        M, N, K = d.M, d.N, d.iters
        off = self.lib.ckpt_layout(d)
        raw = fwd["ckpt"].get()
        n = (K - 1) * N * M
        a = raw[off[3]:off[3] + 4 * n].view(np.float32).reshape(K - 1, N, M)
        return [torch.from_numpy(np.ascontiguousarray(a[k].T)) for k in range(K - 1)]



class EmuBackend(Backend):
    gpu = False


class GpuBackend(Backend):
    gpu = True
