"""ctypes binding of the C ABI in ``include/admmtv.h`` (libadmmtv.so, built by nvcc for sm_100a).

There is no CPU fallback: if the shared library is missing, ``load()`` raises, and every compute
entry point of the library itself fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADMMTV_LIB") or os.path.join(_HERE, "libadmmtv.so")   # same override as julia/ADMMTV.jl

ACT = {"identity": 0, "relu": 1, "relu6": 2, "relu1": 3}
FLAG_NO_CLAMP = 1
FLAG_NOGRAD_REPEAT = 2
FLAG_SHARED_INPUT = 4
FLAG_CHANNEL_CONCAT = 8
FLAG_ISO_PRECOMPUTE = 16
FLAG_ISO_INLINE = 32
FLAG_PER_ITER_PARAMS = 64
FLAG_NO_SMALL = 128

# every symbol include/admmtv.h declares (tests check the .so exports exactly these)
SYMBOLS = (
    "admmtv_version",
    "admmtv_strerror",
    "admmtv_check",
    "admmtv_workspace_bytes",
    "admmtv_forward",
    "admmtv_backward",
    "admmtv_forward_host",
    "admmtv_profile_forward",
    "admmtv_profile_backward",
    "admmtv_ckpt_layout",
    "admmtv_forward_launches",
    "admmtv_backward_launches",
    "admmtv_forward_ex",
    "admmtv_backward_ex",
    "admmtv_backward_mse",
)
# every symbol include/admmtv_loss.h declares
LOSS_SYMBOLS = (
    "admmtv_gmsd_workspace_bytes",
    "admmtv_gmsd_forward",
    "admmtv_gmsd_backward",
    "admmtv_ssim_workspace_bytes",
    "admmtv_ssim_forward",
    "admmtv_ssim_backward",
    "admmtv_ssim_window_workspace_bytes",
    "admmtv_ssim_window_forward",
    "admmtv_ssim_window_backward",
    "admmtv_pad_symmetric",
    "admmtv_pad_symmetric_adjoint",
)
# every symbol include/admmtv_host.h declares
HOST_SYMBOLS = (
    "admmtv_host_session_bytes",
    "admmtv_host_session_create",
    "admmtv_host_session_destroy",
    "admmtv_host_pin",
    "admmtv_host_unpin",
    "admmtv_host_forward_enqueue",
    "admmtv_host_forward_enqueue_n0f8",
    "admmtv_host_train_step_enqueue",
    "admmtv_host_train_step_enqueue_n0f8",
    "admmtv_host_grad_floats",
    "admmtv_host_wait",
    "admmtv_host_launches",
    "admmtv_mse_train_step",
)
# every symbol include/admmtv_batch.h declares
BATCH_SYMBOLS = ("admmtv_batch_from_n0f8", "admmtv_batch_gather_n0f8")


class Desc(C.Structure):
    """struct admmtv_desc (include/admmtv.h)."""

    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("P", C.c_int32), ("B", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32),
        ("iters", C.c_int32), ("iso", C.c_int32), ("activation", C.c_int32), ("has_bias", C.c_int32),
        ("device", C.c_int32), ("flags", C.c_int32),
        ("creg", C.c_float), ("groups", C.c_int32),
    ]


# int (*admmtv_allreduce_fn)(float* buf, size_t count, void* stream, void* user)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)


class Hooks(C.Structure):
    """struct admmtv_hooks (include/admmtv.h)."""

    _fields_ = [("allreduce_sum", ALLREDUCE_FN), ("user", C.c_void_p), ("tau_owner", C.c_int32)]


class AdmmTvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{msg} (code {code})")
        self.code = code


class AdmmTvLib:
    """Thin typed wrapper; pointers are plain integers (device or host addresses)."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise OSError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for the ADMM-TV path."
            )
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        vp, sz = C.c_void_p, C.c_size_t
        L.admmtv_version.restype = C.c_int
        L.admmtv_strerror.restype = C.c_char_p
        L.admmtv_strerror.argtypes = [C.c_int]
        L.admmtv_check.argtypes = [C.POINTER(Desc)]
        L.admmtv_workspace_bytes.argtypes = [C.POINTER(Desc), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
        L.admmtv_forward.argtypes = [C.POINTER(Desc)] + [vp] * 9
        L.admmtv_backward.argtypes = [C.POINTER(Desc)] + [vp] * 14
        L.admmtv_forward_ex.argtypes = [C.POINTER(Desc)] + [vp] * 9 + [C.POINTER(Hooks)]
        L.admmtv_backward_ex.argtypes = [C.POINTER(Desc)] + [vp] * 14 + [C.POINTER(Hooks)]
        L.admmtv_backward_mse.argtypes = [C.POINTER(Desc)] + [vp] * 15
        L.admmtv_profile_forward.argtypes = [C.POINTER(Desc)] + [vp] * 10
        L.admmtv_profile_backward.argtypes = [C.POINTER(Desc)] + [vp] * 15
        L.admmtv_forward_host.argtypes = [C.POINTER(Desc)] + [vp] * 6
        L.admmtv_ckpt_layout.argtypes = [C.POINTER(Desc), C.POINTER(sz)]
        L.admmtv_forward_launches.argtypes = [C.POINTER(Desc), C.c_int]
        L.admmtv_backward_launches.argtypes = [C.POINTER(Desc)]
        i, f = C.c_int, C.c_float
        L.admmtv_gmsd_workspace_bytes.argtypes = [i, i, i, i, C.POINTER(sz)]
        L.admmtv_gmsd_forward.argtypes = [i, i, i, i, i, vp, vp, f, f, vp, vp, vp]
        L.admmtv_gmsd_backward.argtypes = [i, i, i, i, i, vp, vp, f, f, vp, vp, vp, vp]
        L.admmtv_ssim_workspace_bytes.argtypes = [i, i, i, i, i, i, C.POINTER(sz)]
        L.admmtv_ssim_forward.argtypes = [i, i, i, i, i, vp, vp, C.POINTER(f), i, f, i, vp, vp, i, vp]
        L.admmtv_ssim_backward.argtypes = [i, i, i, i, i, vp, vp, C.POINTER(f), i, i, vp, vp, vp, vp]
        fp = C.POINTER(f)
        L.admmtv_ssim_window_workspace_bytes.argtypes = [i, i, i, i, i, i, i, C.POINTER(sz)]
        L.admmtv_ssim_window_forward.argtypes = [i, i, i, i, i, vp, vp, fp, fp, i, i, i, f, i, vp, vp, i, vp]
        L.admmtv_ssim_window_backward.argtypes = [i, i, i, i, i, vp, vp, fp, fp, i, i, i, i, vp, vp, vp, vp]
        L.admmtv_pad_symmetric.argtypes = [i, i, i, i, i, i, i, i, vp, vp, vp]
        L.admmtv_pad_symmetric_adjoint.argtypes = [i, i, i, i, i, i, i, i, vp, vp, vp]
        L.admmtv_batch_from_n0f8.argtypes = [i, i, i, i, i, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, vp]
        L.admmtv_batch_gather_n0f8.argtypes = [i, i, i, i, i, vp, vp, C.c_int64, C.c_int64, C.c_int64, vp, vp]
        self.has_host = hasattr(L, "admmtv_host_wait")   # the test-only CPU emulation build has no host-buffer layer
        if self.has_host:
            L.admmtv_host_session_bytes.argtypes = [C.POINTER(Desc), i, C.POINTER(sz)]
            L.admmtv_host_session_create.argtypes = [C.POINTER(Desc), i, vp, vp, C.POINTER(vp)]
            L.admmtv_host_session_destroy.argtypes = [vp]
            L.admmtv_host_pin.argtypes = [vp, sz]
            L.admmtv_host_unpin.argtypes = [vp]
            L.admmtv_host_forward_enqueue.argtypes = [vp, i] + [vp] * 6
            L.admmtv_host_train_step_enqueue.argtypes = [vp, i] + [vp] * 9 + [C.POINTER(Hooks)]
            L.admmtv_host_forward_enqueue_n0f8.argtypes = [vp, i, vp] + [C.c_int64] * 4 + [vp] * 5
            L.admmtv_host_train_step_enqueue_n0f8.argtypes = [vp, i, vp, vp] + [C.c_int64] * 4 + [vp] * 6 + [C.POINTER(Hooks)]
            L.admmtv_host_grad_floats.argtypes = [C.POINTER(Desc)]
            L.admmtv_host_wait.argtypes = [vp, i]
            L.admmtv_host_launches.argtypes = [vp, i]
            L.admmtv_mse_train_step.argtypes = [C.POINTER(Desc)] + [vp] * 14 + [C.POINTER(Hooks)]
        for name in SYMBOLS + LOSS_SYMBOLS + BATCH_SYMBOLS + (HOST_SYMBOLS if self.has_host else ()):
            getattr(L, name)  # AttributeError if a declared symbol is not exported

    def strerror(self, code: int) -> str:
        return self.lib.admmtv_strerror(code).decode()

    def _raise(self, code: int):
        if code != 0:
            raise AdmmTvError(code, self.strerror(code))

    def version(self) -> int:
        return self.lib.admmtv_version()

    def check(self, d: Desc) -> int:
        return self.lib.admmtv_check(C.byref(d))

    def workspace_bytes(self, d: Desc):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._raise(self.lib.admmtv_workspace_bytes(C.byref(d), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def forward(self, d: Desc, y, h, lam, rho, bias, x_out, ws, ckpt, stream=0):
        self._raise(self.lib.admmtv_forward(C.byref(d), y, h, lam, rho, bias, x_out, ws, ckpt, stream))

    def backward(self, d: Desc, xbar, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar, biasbar, ws, stream=0):
        self._raise(self.lib.admmtv_backward(C.byref(d), xbar, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar,
                                             biasbar, ws, stream))

    def backward_mse(self, d: Desc, target, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar, biasbar, loss_sum, ws, stream=0):
        self._raise(self.lib.admmtv_backward_mse(C.byref(d), target, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar,
                                                 biasbar, loss_sum, ws, stream))

    def forward_ex(self, d: Desc, y, h, lam, rho, bias, x_out, ws, ckpt, stream, hooks: "Hooks"):
        self._raise(self.lib.admmtv_forward_ex(C.byref(d), y, h, lam, rho, bias, x_out, ws, ckpt, stream, C.byref(hooks)))

    def backward_ex(self, d: Desc, xbar, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar, biasbar, ws, stream,
                    hooks: "Hooks"):
        self._raise(self.lib.admmtv_backward_ex(C.byref(d), xbar, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar,
                                                biasbar, ws, stream, C.byref(hooks)))

    def profile_forward(self, d: Desc, y, h, lam, rho, bias, x_out, ws, ckpt, stream=0):
        """Returns (total_ms, dim2_ms, dim1_ms, other_ms); synchronises."""
        ms = (C.c_float * 4)()
        self._raise(self.lib.admmtv_profile_forward(C.byref(d), y, h, lam, rho, bias, x_out, ws, ckpt, stream, ms))
        return tuple(ms)

    def profile_backward(self, d: Desc, xbar, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar, biasbar, ws, stream=0):
        ms = (C.c_float * 4)()
        self._raise(self.lib.admmtv_profile_backward(C.byref(d), xbar, x_out, y, h, lam, rho, ckpt, ybar, hbar, lambar, rhobar,
                                                     biasbar, ws, stream, ms))
        return tuple(ms)

    def forward_host(self, d: Desc, y, h, lam, rho, bias, x_out):
        self._raise(self.lib.admmtv_forward_host(C.byref(d), y, h, lam, rho, bias, x_out))

    def ckpt_layout(self, d: Desc):
        out = (C.c_size_t * 4)()
        self._raise(self.lib.admmtv_ckpt_layout(C.byref(d), out))
        return tuple(out)

    def forward_launches(self, d: Desc, with_ckpt: bool) -> int:
        return self.lib.admmtv_forward_launches(C.byref(d), int(with_ckpt))

    def backward_launches(self, d: Desc) -> int:
        return self.lib.admmtv_backward_launches(C.byref(d))

    # ---- include/admmtv_host.h ----------------------------------------------------------------
    def host_session_bytes(self, d: Desc, training: bool) -> int:
        n = C.c_size_t()
        self._raise(self.lib.admmtv_host_session_bytes(C.byref(d), int(training), C.byref(n)))
        return n.value

    def host_session_create(self, d: Desc, training: bool, arena=None, compute_stream=None) -> int:
        h = C.c_void_p()
        self._raise(self.lib.admmtv_host_session_create(C.byref(d), int(training), arena, compute_stream, C.byref(h)))
        return h.value

    def host_session_destroy(self, sess):
        self._raise(self.lib.admmtv_host_session_destroy(sess))

    def host_pin(self, ptr, nbytes):
        self._raise(self.lib.admmtv_host_pin(ptr, nbytes))

    def host_unpin(self, ptr):
        self._raise(self.lib.admmtv_host_unpin(ptr))

    def host_forward_enqueue(self, sess, slot, y, h, lam, rho, bias, x_out):
        self._raise(self.lib.admmtv_host_forward_enqueue(sess, slot, y, h, lam, rho, bias, x_out))

    def host_train_step_enqueue(self, sess, slot, y, target, h, lam, rho, bias, grads_out, loss_out, ybar_out=None, hooks=None):
        self._raise(self.lib.admmtv_host_train_step_enqueue(sess, slot, y, target, h, lam, rho, bias, grads_out, loss_out, ybar_out,
                                                            None if hooks is None else C.byref(hooks)))

    def host_forward_enqueue_n0f8(self, sess, slot, y, strides, h, lam, rho, bias, x_out):
        sc, si, sj, sb = strides
        self._raise(self.lib.admmtv_host_forward_enqueue_n0f8(sess, slot, y, sc, si, sj, sb, h, lam, rho, bias, x_out))

    def host_train_step_enqueue_n0f8(self, sess, slot, y, target, strides, h, lam, rho, bias, grads_out, loss_out, hooks=None):
        sc, si, sj, sb = strides
        self._raise(self.lib.admmtv_host_train_step_enqueue_n0f8(sess, slot, y, target, sc, si, sj, sb, h, lam, rho, bias, grads_out,
                                                                 loss_out, None if hooks is None else C.byref(hooks)))

    def mse_train_step(self, d: Desc, y, target, h, lam, rho, bias, x_out, ybar, grads, loss_sum, ws_fwd, ckpt, ws_bwd,
                       stream=0, hooks=None):
        self._raise(self.lib.admmtv_mse_train_step(C.byref(d), y, target, h, lam, rho, bias, x_out, ybar, grads, loss_sum,
                                                   ws_fwd, ckpt, ws_bwd, stream, None if hooks is None else C.byref(hooks)))

    def host_grad_floats(self, d: Desc) -> int:
        return self.lib.admmtv_host_grad_floats(C.byref(d))

    def host_wait(self, sess, slot):
        self._raise(self.lib.admmtv_host_wait(sess, slot))

    def host_launches(self, sess, training: bool) -> int:
        return self.lib.admmtv_host_launches(sess, int(training))

    # ---- include/admmtv_loss.h ----------------------------------------------------------------
    @staticmethod
    def _taps(taps):
        if taps is None:
            return None, 0
        arr = (C.c_float * len(taps))(*[float(t) for t in taps])
        return arr, len(taps)

    def gmsd_workspace_bytes(self, M, N, Cc, B) -> int:
        n = C.c_size_t()
        self._raise(self.lib.admmtv_gmsd_workspace_bytes(M, N, Cc, B, C.byref(n)))
        return n.value

    def gmsd_forward(self, M, N, Cc, B, device, x, y, t, alpha, loss_out, ws, stream=0):
        self._raise(self.lib.admmtv_gmsd_forward(M, N, Cc, B, device, x, y, t, alpha, loss_out, ws, stream))

    def gmsd_backward(self, M, N, Cc, B, device, x, y, t, alpha, lossbar, ws, xbar, stream=0):
        self._raise(self.lib.admmtv_gmsd_backward(M, N, Cc, B, device, x, y, t, alpha, lossbar, ws, xbar, stream))

    def ssim_workspace_bytes(self, M, N, Cc, B, taps=None, with_grad=True) -> int:
        n = C.c_size_t()
        self._raise(self.lib.admmtv_ssim_workspace_bytes(M, N, Cc, B, 0 if taps is None else len(taps), int(with_grad), C.byref(n)))
        return n.value

    def ssim_forward(self, M, N, Cc, B, device, x, y, taps, peakval, as_loss, out, ws, with_grad, stream=0):
        arr, L = self._taps(taps)
        self._raise(self.lib.admmtv_ssim_forward(M, N, Cc, B, device, x, y, arr, L, peakval, int(as_loss), out, ws,
                                                 int(with_grad), stream))

    @staticmethod
    def _factors(u, v):
        """u: R x L1, v: R x L2 nested sequences -> (float arrays, L1, L2, R)"""
        R, L1, L2 = len(u), len(u[0]), len(v[0])
        ua = (C.c_float * (R * L1))(*[float(t) for row in u for t in row])
        va = (C.c_float * (R * L2))(*[float(t) for row in v for t in row])
        return ua, va, L1, L2, R

    def ssim_window_workspace_bytes(self, M, N, Cc, B, L1, L2, with_grad=True) -> int:
        n = C.c_size_t()
        self._raise(self.lib.admmtv_ssim_window_workspace_bytes(M, N, Cc, B, L1, L2, int(with_grad), C.byref(n)))
        return n.value

    def ssim_window_forward(self, M, N, Cc, B, device, x, y, u, v, peakval, as_loss, out, ws, with_grad, stream=0):
        ua, va, L1, L2, R = self._factors(u, v)
        self._raise(self.lib.admmtv_ssim_window_forward(M, N, Cc, B, device, x, y, ua, va, L1, L2, R, peakval, int(as_loss),
                                                        out, ws, int(with_grad), stream))

    def ssim_window_backward(self, M, N, Cc, B, device, x, y, u, v, as_loss, outbar, ws, xbar, stream=0):
        ua, va, L1, L2, R = self._factors(u, v)
        self._raise(self.lib.admmtv_ssim_window_backward(M, N, Cc, B, device, x, y, ua, va, L1, L2, R, int(as_loss), outbar,
                                                         ws, xbar, stream))

    def pad_symmetric(self, M, N, planes, pads, device, src, dst, stream=0):
        lo1, hi1, lo2, hi2 = pads
        self._raise(self.lib.admmtv_pad_symmetric(M, N, planes, lo1, hi1, lo2, hi2, device, src, dst, stream))

    def pad_symmetric_adjoint(self, M, N, planes, pads, device, padded_bar, src_bar, stream=0):
        lo1, hi1, lo2, hi2 = pads
        self._raise(self.lib.admmtv_pad_symmetric_adjoint(M, N, planes, lo1, hi1, lo2, hi2, device, padded_bar, src_bar, stream))

    def batch_from_n0f8(self, M, N, Cc, B, device, src, sc, si, sj, sb, dst, stream=0):
        self._raise(self.lib.admmtv_batch_from_n0f8(M, N, Cc, B, device, src, sc, si, sj, sb, dst, stream))

    def batch_gather_n0f8(self, M, N, Cc, B, device, base, offsets, sc, si, sj, dst, stream=0):
        self._raise(self.lib.admmtv_batch_gather_n0f8(M, N, Cc, B, device, base, offsets, sc, si, sj, dst, stream))

    def ssim_backward(self, M, N, Cc, B, device, x, y, taps, as_loss, outbar, ws, xbar, stream=0):
        arr, L = self._taps(taps)
        self._raise(self.lib.admmtv_ssim_backward(M, N, Cc, B, device, x, y, arr, L, int(as_loss), outbar, ws, xbar, stream))


_LIB: Optional[AdmmTvLib] = None


def load() -> AdmmTvLib:
    """The product library.  Raises if it has not been built."""
    global _LIB
    if _LIB is None:
        _LIB = AdmmTvLib(LIB_PATH)
    return _LIB


def make_desc(M, N, P, B, kh, kw, iters, iso=False, activation="identity", has_bias=False, device=0, flags=0,
              creg=0.0, groups=0) -> Desc:
    act = ACT[activation] if isinstance(activation, str) else int(activation)
    return Desc(M, N, P, B, kh, kw, iters, int(bool(iso)), act, int(bool(has_bias)), device, flags, float(creg), int(groups))
