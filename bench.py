#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (task contract).

    python bench.py --gpus 1 --steps K --warmup W            native arm (CUDA kernels through the C ABI)
    python bench.py --impl reference ...                      reference arm (restated CPU path, same config)

Headline ("value"): BASELINE.json's metric -- ADMM megapixel-iterations/s, forward + backward + gradient all-reduce --
on configs[1]'s batch (64 x 512x512 RGB, motion-blur PSF 15x15) with 10 unrolled iterations per step, inputs resident in
HBM (admmtv_mse_train_step, include/admmtv_host.h).  "e2e" is the same step through the host-buffer C-ABI session
(admmtv_host_train_step_enqueue_n0f8 / _wait): batch and target uploaded from pinned host memory as the 8-bit samples
they are, the gradients + loss downloaded, inside the timed region, every step ("e2e_f32": the same values as fp32 arrays).  The "others" block carries short runs of every other BASELINE config
(configs[1] forward-only with 100 iterations, configs[2] strong-scaled over the ranks, configs[3], configs[4] with
per-image PSFs), each with its own time, roofline fraction and clock record.  Weak scaling: every rank gets its own
batch of 64 for the headline.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "admm_tv_plane_megapixel_iterations_per_second"
UNIT = "Mpx-it/s"

WORKLOADS = {
    "cfg2_train": dict(B=64, P=3, N=512, M=512, k=15, iters=10, mode="fwd+bwd", psf="motion",
                       desc="BASELINE.json configs[1] batch (64 x 512x512 RGB, motion-blur PSF 15x15), 10 unrolled ADMM-TV "
                            "iterations, forward (checkpointed) + hand-written backward + gradient all-reduce, MSE loss"),
    "cfg2": dict(B=64, P=3, N=512, M=512, k=15, iters=100, mode="fwd", psf="motion",
                 desc="BASELINE.json configs[1]: batch 64 x 512x512 RGB, motion-blur PSF 15x15, ADMM-TV forward 100 iterations"),
    "cfg3": dict(B=256, P=3, N=256, M=256, k=15, iters=10, mode="fwd+bwd", psf="gauss", strong=True,
                 desc="BASELINE.json configs[2]: training step, 10 learned-rho/lambda iterations, fwd+bwd, GLOBAL batch "
                      "256 x 256x256 RGB split over the ranks (strong scaling), Gaussian PSF 15x15, NCCL gradient all-reduce"),
    "cfg3_share": dict(B=32, P=3, N=256, M=256, k=15, iters=10, mode="fwd+bwd", psf="gauss",
                       desc="one GPU's share of BASELINE.json configs[2] at 8 GPUs: 32 x 256x256 RGB, 10 iterations, fwd+bwd"),
    "cfg4": dict(B=16, P=1, N=2048, M=2048, k=31, iters=200, mode="fwd", psf="motion",
                 desc="BASELINE.json configs[3]: batch 16 x 2048x2048 gray, 31x31 PSF, 200-iteration forward"),
    "w4096": dict(B=4, P=1, N=4096, M=4096, k=31, iters=50, mode="fwd", psf="motion",
                  desc="largest planned FFT length: batch 4 x 4096x4096 gray, 31x31 PSF, 50-iteration forward (tuning workload)"),
    "cfg5": dict(B=1024, P=1, N=128, M=128, k=9, iters=50, mode="fwd", psf="motion",
                 desc="BASELINE.json configs[4], shared-PSF variant: batch 1024 x 128x128, 50 iterations"),
    "cfg5_mixed": dict(B=1024, P=1, N=128, M=128, k=9, iters=50, mode="grouped", groups="per_image", psf="motion",
                       desc="BASELINE.json configs[4]: batch 1024 x 128x128, PER-IMAGE motion PSFs 9x9 and noise levels "
                            "(lambda_i = 0.2 sigma_i, rho_i = 5 lambda_i), 50 iterations, one grouped call"),
    "cfg2_iso": dict(B=64, P=3, N=512, M=512, k=15, iters=50, mode="fwd", psf="motion", iso=True,
                     desc="configs[1] batch, ISOTROPIC TV (the reference's shipped use_iso = true), 50 iterations forward"),
    "denoiser5": dict(B=32, P=3, N=256, M=256, k=0, iters=50, mode="grouped", groups=5, iso=True, psf="motion",
                      desc="net_build.jl:113-128 get_denoiser: 5 parallel ADMMDeconvF2((),50,rho_i,relu1; iso) branches on the "
                           "same 32 x 256x256 RGB input, channel-concatenated, one grouped call"),
    "hd1080": dict(B=4, P=3, N=1920, M=1080, k=15, iters=20, mode="fwd", psf="motion",
                   desc="4 x 1080x1920 RGB frames (1080 has no register-FFT plan: generic-size kernels), motion PSF 15x15, 20 iterations"),
    "bsd481": dict(B=32, P=3, N=481, M=321, k=9, iters=20, mode="fwd", psf="motion",
                   desc="32 x 321x481 RGB (BSD500 size; 481 = 13*37: generic-size kernels), motion PSF 9x9, 20 iterations"),
    "vga": dict(B=64, P=3, N=480, M=640, k=15, iters=50, mode="fwd", psf="motion",
                desc="64 x 640x480 RGB frames (mixed-radix lengths 640 = 5*16*8, 480 = 3*5*8*4), motion PSF 15x15, 50 iterations"),
    "tiny": dict(B=2, P=3, N=64, M=64, k=7, iters=4, mode="fwd+bwd", psf="motion", desc="tiny debug workload"),
}
OTHERS = ("cfg2", "cfg3", "cfg4", "cfg5_mixed", "cfg2_iso")
FWD_BYTES = 40.0   # algorithmic bytes / plane-pixel-iteration, forward  
BWD_BYTES = 68.0
ISO_FWD_BYTES = 44.0   # + the per-pixel norm / scale


def alg_bytes(w):
    if w["mode"] == "fwd+bwd":
        return FWD_BYTES + BWD_BYTES
    return ISO_FWD_BYTES if w.get("iso") else FWD_BYTES


def config_of(w, name):
    """The `config` object of the JSON line: identical on the native and the reference arm."""
    return {"workload": w["desc"], "name": name, "mode": w["mode"], "iters": w["iters"],
            "image": f"{w['M']}x{w['N']}x{w['P']}", "per_gpu_batch": w["B"], "psf": f"{w['k']}x{w['k']} {w['psf']}",
            "lambda": 0.0041, "rho": 0.021, "iso": bool(w.get("iso", False)), "loss": "mse" if w["mode"] == "fwd+bwd" else None,
            "samples": "8-bit image values k/255 held as fp32" if w["mode"] == "fwd+bwd" else "fp32",
            "l2": "per-iteration state is larger than the 126 MB L2 (no flush needed)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling during a timed region (the exact PID is terminated afterwards)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True, period_ms: int = 50):
        self.index, self.proc, self.enabled, self.period = index, None, enabled, period_ms

    def start(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period), "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        if self.proc is not None:
            self.first = self.proc.stdout.readline()      # block until nvidia-smi is up and has delivered its first sample
        return self

    def stop(self):
        if not self.enabled:
            return None
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        out = getattr(self, "first", "") + out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def motion_psf(k: int, theta: float, length: float):
    """Linear motion blur through the centre of a k x k support, bilinear rasterised, sum 1.
    Returns a (k,k) torch tensor indexed [dim1, dim2]."""
    import torch
    p = torch.zeros(k, k, dtype=torch.float64)
    c = (k - 1) / 2
    n = max(int(math.ceil(length * 4)), 2)
    for t in range(n):
        s = -length / 2 + length * t / (n - 1)
        a, b = c + s * math.sin(theta), c + s * math.cos(theta)
        a0, b0 = int(math.floor(a)), int(math.floor(b))
        for da in (0, 1):
            for db in (0, 1):
                aa, bb = a0 + da, b0 + db
                if 0 <= aa < k and 0 <= bb < k:
                    p[aa, bb] += (1 - abs(a - aa)) * (1 - abs(b - bb))
    return (p / p.sum()).float()


def gauss_psf(k: int, sigma: float):
    import torch
    x = torch.arange(k, dtype=torch.float64) - (k - 1) / 2
    g = torch.exp(-x * x / (2 * sigma * sigma))
    p = g[:, None] * g[None, :]
    return (p / p.sum()).float()


def make_inputs(w, seed, B=None):
    """Synthetic blurred-noisy batch in the (B,P,N,M) layout, its ground truth, and the PSF.  Plain torch;
    nothing from oracle/ is used on the native arm.  Returns (y, g, h (1,1,kw,kh))."""
    import numpy as np
    import torch

    rng = np.random.Generator(np.random.PCG64(seed))
    k = w["k"] if w["k"] > 0 else 9      # k = 0: the layer has no PSF (empty weight); the scene is still blurred
    if w.get("psf") == "gauss":
        h = gauss_psf(k, 2.0)
    else:
        h = motion_psf(k, float(rng.uniform(0, math.pi)), float(rng.uniform(5, k)))      # [dim1, dim2]
    B = w["B"] if B is None else B
    P, N, M = w["P"], w["N"], w["M"]
    nb = min(B, 4)
    g = torch.from_numpy(rng.random((nb, P, N, M), dtype=np.float32))
    g = torch.nn.functional.avg_pool2d(g, 9, stride=1, padding=4, count_include_pad=False)
    for b in range(nb):
        for _ in range(8):
            i0, i1 = sorted(rng.integers(0, N, 2).tolist()); j0, j1 = sorted(rng.integers(0, M, 2).tolist())
            g[b, :, i0:i1 + 1, j0:j1 + 1] = float(rng.random())
    # circular blur with the reference's alignment (true convolution, centre at ceil((k-1)/2)): in the (.., N, M)
    # layout dim 1 (M) is the last axis, dim 2 (N) the one before
    pu, pd = math.ceil((k - 1) / 2), (k - 1) // 2
    hk = h.t().contiguous()                                            # [dim2, dim1] == (kw, kh)
    gp = torch.nn.functional.pad(g.reshape(-1, 1, N, M), (pu, pd, pu, pd), mode="circular")
    y = torch.nn.functional.conv2d(gp, torch.flip(hk, dims=(0, 1)).reshape(1, 1, k, k)).reshape(g.shape)
    y = y + 0.02 * torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    if w["mode"] == "fwd+bwd":
        # training batches are decoded 8-bit images (datafeeder.jl:54-68): both arms, the resident step and the host-buffer
        # step see the same values k / 255, and the host-buffer step can upload them as bytes (e2e below)
        y = (y.clamp(0, 1) * 255).round() / 255
        g = (g.clamp(0, 1) * 255).round() / 255
    reps = (B + nb - 1) // nb
    y = y.repeat(reps, 1, 1, 1)[:B].contiguous()
    g = g.repeat(reps, 1, 1, 1)[:B].contiguous()
    return y, g, hk.reshape(1, 1, k, k).contiguous()                      # (1,1,kw,kh)


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the restated reference (oracle port of ops.jl:17-96, torch-CPU fp32, MKL FFT) -- the only place
# bench.py touches oracle/.  fwd+bwd = torch.autograd through the literal restatement (stands in for Zygote).
# ---------------------------------------------------------------------------------------------------------------
def cpu_step_fn(w, sB):
    import torch
    from oracle import admm_tv_oracle as O

    y, g, hk = make_inputs(w, 1001, B=sB)
    yj = y.permute(3, 2, 1, 0).contiguous(); gj = g.permute(3, 2, 1, 0).contiguous()
    hj = hk.permute(3, 2, 1, 0).contiguous()
    K, iso = w["iters"], bool(w.get("iso", False))
    if w["mode"] == "fwd+bwd":
        def step():
            lam = torch.tensor([0.0041], requires_grad=True); rho = torch.tensor([0.021], requires_grad=True)
            h = hj.clone().requires_grad_(True)
            x = O.tvd_fft_cpu(yj, lam, rho, h, iso, K)
            loss = ((x - gj) ** 2).mean()
            loss.backward()
            return float(loss.detach())
    else:
        lam = torch.tensor([0.0041]); rho = torch.tensor([0.021])

        def step():
            with torch.no_grad():
                O.tvd_fft_cpu(yj, lam, rho, hj, iso, K)
    units = sB * w["P"] * w["N"] * w["M"] * K / 1e6
    what = "forward + torch.autograd backward" if w["mode"] == "fwd+bwd" else "forward"
    sample = (f"{sB} x {w['M']}x{w['N']}x{w['P']} images of the workload, all {K} iterations, {what}, torch-CPU fp32 (MKL FFT), "
              "faithful (H^T y recomputed every iteration as ops.jl:86 does)")
    return step, units, sample


def run_reference(args, w, name):
    """Reference arm: the restated reference CPU path on every host thread, same config, bounded sample per step."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sB = 2 if w["M"] * w["N"] <= 512 * 512 else 1
    step, units, sample = cpu_step_fn(w, sB)
    for _ in range(max(min(args.warmup, 2), 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = units * args.steps / dt
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(w, name),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated reference (oracle/admm_tv_oracle.py, torch-CPU fp32, MKL FFT); Julia/FFTW cannot run in this image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def cpu_baseline(w):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sB = 2 if w["M"] * w["N"] <= 512 * 512 else 1
    step, units, sample = cpu_step_fn(w, sB)
    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step(); n += 1
        dt = time.perf_counter() - t0
        if dt > 12.0 or n >= 20:
            break
    return {"value": units * n / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + f"; {n} repetitions"}


# ---------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def setup_dist():
    import torch
    import torch.distributed as dist

    c = Ctx()
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if c.world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{c.local}"))
    torch.cuda.set_device(c.local)
    if c.world > 1:
        # one process per GPU: run (and first-touch the pinned host buffers) on the CPUs of the GPU's own NUMA node, so that the
        # e2e uploads of eight ranks do not all cross the socket interconnect.  Not done at N = 1, where rank 0 also runs the
        # CPU baseline on every host core.
        try:
            import pynvml
            pynvml.nvmlInit()
            try:      # NVML enumerates physical devices: match by UUID when the runtime exposes it
                hnd = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(c.local).uuid)).encode())
            except Exception:
                hnd = pynvml.nvmlDeviceGetHandleByIndex(c.local)
            pynvml.nvmlDeviceSetCpuAffinity(hnd)
            c.affinity = len(os.sched_getaffinity(0))
        except Exception:
            c.affinity = None
    c.dev = torch.device(f"cuda:{c.local}")
    c.dist = dist
    return c


def barrier(c):
    import torch
    if c.world > 1:
        c.dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(c, vals):
    import torch
    t = torch.tensor(vals, device=c.dev, dtype=torch.float64)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return [float(v) for v in t]


def timed(c, fn, steps, warmup, clocks=True, min_ms=0.0):
    """W warm-up calls, then `steps` calls bracketed by barrier + synchronize, CUDA events on the launching stream;
    returns (ms per step as the max over ranks, clock record of rank 0, steps timed).  min_ms > 0 raises the number of
    timed steps so that the timed region lasts at least that long (enough nvidia-smi clock samples for short workloads)."""
    import torch
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nw = max(warmup, 3)
    fn()
    w0.record()
    for _ in range(nw - 1):
        fn()
    w1.record()
    barrier(c)
    if min_ms > 0:
        est = max(w0.elapsed_time(w1) / max(nw - 1, 1), 1e-3)
        steps = int(max_over_ranks(c, [max(steps, math.ceil(min_ms / est))])[0])
    sampler = ClockSampler(c.local, enabled=clocks and c.rank == 0).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(c)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier(c)
    ms = e0.elapsed_time(e1) / steps
    clk = sampler.stop()
    return max_over_ranks(c, [ms])[0], clk, steps


class TrainStep:
    """forward (checkpointed) + MSE + backward + gradient all-reduce on resident inputs: admmtv_mse_train_step."""

    def __init__(self, c, w, B, seed, graph=False):
        import torch
        self.graph = None
        self.want_graph = graph
        import admm_deconv_b200 as A
        from admm_deconv_b200 import _lib

        self.c, self.w, self.B = c, w, B
        self.lib = A.load()
        y, g, h = make_inputs(w, seed, B=B)
        self.y_host, self.g_host = y.pin_memory(), g.pin_memory()
        self.h_host = h
        dev = c.dev
        self.y, self.g = y.to(dev), g.to(dev)
        self.h = h.to(dev).contiguous()
        self.lam = torch.tensor([0.0041], device=dev); self.rho = torch.tensor([0.021], device=dev)
        k = w["k"]
        self.d = _lib.make_desc(w["M"], w["N"], w["P"], B, k, k, w["iters"], bool(w.get("iso", False)), "identity", False,
                                c.local, _lib.FLAG_NO_CLAMP, 0.0)
        fwd_b, ck_b, bwd_b = self.lib.workspace_bytes(self.d)
        u8 = lambda n: torch.empty(max(n, 256), dtype=torch.uint8, device=dev)
        self.ws_f, self.ck, self.ws_b = u8(fwd_b), u8(ck_b), u8(bwd_b)
        self.x, self.ybar = torch.empty_like(self.y), torch.empty_like(self.y)
        self.xbar = torch.full_like(self.y, 1e-6)     # only for the profiling twin of the backward (the timed step forms it on the fly)
        self.ngrad = self.lib.host_grad_floats(self.d)
        self.grads = torch.zeros(self.ngrad, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=dev)
        self.px = B * w["P"] * w["N"] * w["M"]
        self.launches = self.lib.forward_launches(self.d, True) + self.lib.backward_launches(self.d)

    def _enqueue(self):
        import torch
        st = torch.cuda.current_stream().cuda_stream
        p = lambda t: t.data_ptr()
        self.lib.mse_train_step(self.d, p(self.y), p(self.g), p(self.h), p(self.lam), p(self.rho), None, p(self.x),
                                p(self.ybar), p(self.grads), p(self.loss), p(self.ws_f), p(self.ck), p(self.ws_b), st)

    def capture(self):
        """Launch-bound shares (strong scaling): the step's ~55 stream-ordered launches are captured once in a CUDA graph and
        replayed (the library only enqueues kernels and memsets on the caller's stream; the TMA descriptors are by-value
        kernel parameters).  The all-reduce stays outside the graph."""
        import torch
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._enqueue()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue()

    def __call__(self):
        if self.want_graph and self.graph is None:
            self.capture()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        if self.c.world > 1:
            self.c.dist.all_reduce(self.grads)      # [hbar | lambdabar | rhobar]: ONE packed NCCL all-reduce over NVLink

    def profile(self):
        """Per-kernel-class CUDA-event times of one forward and one backward (profiling twins of the same calls)."""
        import torch
        st = torch.cuda.current_stream().cuda_stream
        p = lambda t: t.data_ptr()
        hb, lb, rb = torch.empty_like(self.h), torch.empty_like(self.lam), torch.empty_like(self.rho)
        f = b = None
        for _ in range(2):
            f = self.lib.profile_forward(self.d, p(self.y), p(self.h), p(self.lam), p(self.rho), None, p(self.x), p(self.ws_f), p(self.ck), st)
            b = self.lib.profile_backward(self.d, p(self.xbar), p(self.x), p(self.y), p(self.h), p(self.lam), p(self.rho), p(self.ck),
                                          p(self.ybar), p(hb), p(lb), p(rb), None, p(self.ws_b), st)
        return f, b


def e2e_train(c, w, B, steps, warmup, ts: TrainStep):
    """The same step through the host-buffer C-ABI session, two slots pipelined, copies inside the timed region every step.
    Primary: the batch and the target are uploaded as the 8-bit samples they are (admmtv_host_train_step_enqueue_n0f8, one
    byte per sample, fp32 batch built on the device) and the gradients + loss come back.  "f32": the same values uploaded as
    fp32 arrays (admmtv_host_train_step_enqueue, what `|> gpu` of train.jl:50 moves): four times the bytes, which saturates
    the host's memory once eight ranks share it."""
    import torch
    from admm_deconv_b200 import host

    k = w["k"]
    s = host.HostSession(w["M"], w["N"], w["P"], B, k, k, iters=w["iters"], iso=bool(w.get("iso", False)), device=c.local,
                         flags=1, training=True)
    lam = torch.tensor([0.0041]).pin_memory(); rho = torch.tensor([0.021]).pin_memory()
    h = ts.h_host.clone().pin_memory()
    grads = [torch.empty(s.ngrad).pin_memory() for _ in range(2)]
    loss = [torch.empty(1).pin_memory() for _ in range(2)]
    yu = (ts.y_host * 255).round().to(torch.uint8).pin_memory()      # exact: the batch holds k / 255
    gu = (ts.g_host * 255).round().to(torch.uint8).pin_memory()
    small = (h.numel() + 2) * 4

    def timed(enqueue):
        def run(n):
            for i in range(n):
                sl = i & 1
                if i >= 2:
                    s.wait(sl)
                enqueue(sl)
            s.wait(0); s.wait(1)
        run(max(warmup, 3))
        barrier(c)
        t0 = time.perf_counter()
        run(steps)
        barrier(c)
        return max_over_ranks(c, [(time.perf_counter() - t0) * 1e3 / steps])[0]

    ms8 = timed(lambda sl: s.train_step_enqueue_n0f8(sl, yu, gu, lam, rho, h, grads=grads[sl], loss=loss[sl]))
    loss8 = float(loss[(steps - 1) & 1])
    out = {"ms_per_step": ms8, "h2d_bytes_per_step": yu.numel() + gu.numel() + small,
           "d2h_bytes_per_step": s.ngrad * 4 + 8, "loss": loss8, "launches_per_step": s.launches() + 2,
           "path": "admmtv_host_train_step_enqueue_n0f8 / admmtv_host_wait (include/admmtv_host.h): batch and target as 8-bit "
                   "samples in pinned host memory, fp32 batch built on the device (2 more launches), 2 slots",
           "timer": "host wall clock around the blocking calls (max over ranks)"}
    try:
        msf = timed(lambda sl: s.train_step_enqueue(sl, ts.y_host, ts.g_host, lam, rho, h, grads=grads[sl], loss=loss[sl]))
        out["f32"] = {"ms_per_step": msf, "h2d_bytes_per_step": (ts.y_host.numel() + ts.g_host.numel()) * 4 + small,
                      "d2h_bytes_per_step": s.ngrad * 4 + 8, "loss": float(loss[(steps - 1) & 1]), "launches_per_step": s.launches(),
                      "path": "admmtv_host_train_step_enqueue: the same values as fp32 host arrays"}
    except Exception as e:   # the extra measurement must never take the headline down
        out["f32"] = {"error": repr(e)}
    s.close()
    return out


def run_fwd(c, w, name, steps, warmup, e2e=True, min_ms=0.0):
    """Forward-only workloads (inference), plain or grouped."""
    import numpy as np
    import torch
    from admm_deconv_b200 import host, ops

    dev = c.dev
    y_host, _, h_host = make_inputs(w, 1001 + c.rank)
    y = y_host.to(dev); h = h_host.to(dev)
    lam = torch.tensor([0.0041], device=dev); rho = torch.tensor([0.021], device=dev)
    K, iso = w["iters"], bool(w.get("iso", False))
    px = w["B"] * w["P"] * w["N"] * w["M"]
    grouped = w["mode"] == "grouped"
    if grouped:
        if w["groups"] == "per_image":
            G = w["B"]
            rng = np.random.Generator(np.random.PCG64(7))
            hs = [motion_psf(w["k"], float(rng.uniform(0, math.pi)), float(rng.uniform(5, w["k"]))).t().contiguous() for _ in range(16)]
            hG = torch.stack([hs[i % 16].reshape(1, w["k"], w["k"]) for i in range(G)]).contiguous().to(dev)  # (G,1,kw,kh)
            sig = torch.tensor([[0.005, 0.01, 0.02, 0.04][i % 4] for i in range(G)], device=dev)
            lamG = (0.2 * sig).contiguous(); rhoG = (5 * lamG).contiguous()
            gkw = dict(groups=G)
        else:
            G = int(w["groups"])
            hG = None
            lamG = torch.full((G,), 0.02, device=dev); rhoG = torch.tensor([0.01 * 3 ** i for i in range(G)], device=dev)
            gkw = dict(groups=G, shared_input=True, channel_concat=True, activation="relu1")
            px *= G

        def step():
            with torch.no_grad():
                return ops.tvd_fft_grouped(y, lamG, rhoG, hG, iso, K, **gkw)
    else:
        def step():
            with torch.no_grad():
                return ops.tvd_fft(y, lam, rho, h, iso, K)

    ms, clk, steps = timed(c, step, steps, warmup, min_ms=min_ms)
    pk, _ = peaks()
    r = {"config": config_of(w, name), "n_gpus": c.world, "scaling": "weak", "steps": steps, "ms_per_step": ms,
         "value": px * K * c.world / (ms * 1e-3) / 1e6, "unit": UNIT,
         "algorithmic_bytes_per_plane_pixel_iteration": alg_bytes(w),
         "frac": alg_bytes(w) * px * K / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "clocks": clk}
    if e2e and not grouped:
        k = w["k"]
        s = host.HostSession(w["M"], w["N"], w["P"], w["B"], k, k, iters=K, iso=iso, device=c.local, flags=1)
        yh = y_host.pin_memory(); xo = [torch.empty_like(y_host).pin_memory() for _ in range(2)]
        lh = torch.tensor([0.0041]).pin_memory(); rh = torch.tensor([0.021]).pin_memory(); hh = h_host.clone().pin_memory()

        def run(n):
            for i in range(n):
                sl = i & 1
                if i >= 2:
                    s.wait(sl)
                s.forward_enqueue(sl, yh, lh, rh, hh, out=xo[sl])
            s.wait(0); s.wait(1)

        run(2)
        barrier(c)
        t0 = time.perf_counter()
        run(steps)
        barrier(c)
        ms_e = max_over_ranks(c, [(time.perf_counter() - t0) * 1e3 / steps])[0]
        r["e2e"] = {"value": px * K * c.world / (ms_e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e,
                    "h2d_bytes_per_step": yh.numel() * 4, "d2h_bytes_per_step": yh.numel() * 4,
                    "path": "admmtv_host_forward_enqueue / admmtv_host_wait, pinned host buffers, 2 slots"}
        s.close()
    return r


args_graph = None   # --graph 0/1 overrides the default (CUDA-graph replay of the step when the per-GPU share is below 16 Mpx)


def run_train(c, w, name, steps, warmup, full, min_ms=0.0):
    """fwd+bwd workloads.  full: roofline + e2e + launch count (the headline); else a short `others` entry."""
    import torch

    B = w["B"] // c.world if w.get("strong") else w["B"]
    use_graph = bool(args_graph) if args_graph is not None else (B * w["P"] * w["N"] * w["M"] < (16 << 20))
    ts = TrainStep(c, w, B, 1001 + c.rank, graph=use_graph)
    ms, clk, steps = timed(c, ts, steps, warmup, min_ms=min_ms)
    pk, pk_src = peaks()
    K = w["iters"]
    px_all = ts.px * c.world
    r = {"config": config_of(w, name), "n_gpus": c.world, "scaling": "strong" if w.get("strong") else "weak", "steps": steps,
         "ms_per_step": ms, "value": px_all * K / (ms * 1e-3) / 1e6, "unit": UNIT,
         "algorithmic_bytes_per_plane_pixel_iteration": alg_bytes(w),
         "frac": alg_bytes(w) * ts.px * K / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "clocks": clk,
         "per_gpu_batch": B, "cuda_graph": use_graph, "collective": "one packed NCCL all-reduce of %d floats per step" % ts.ngrad if c.world > 1 else None}
    r["config"]["per_gpu_batch"] = B
    if not full:
        del ts
        torch.cuda.empty_cache()
        return r, None
    f, b = ts.profile()
    n2, n1 = K, max(K - 1, 1)
    kt = {"k_dim2t<save F r_k>": f[1] / n2, "k_dim1_fwd_tma": f[2] / n1, "k_dim2<G += Re(conj Z Z2)>": b[1] / n2, "k_dim1_bwd_tma": b[2] / n1}
    actual = {"k_dim2t<save F r_k>": 12.0, "k_dim1_fwd_tma": 28.0, "k_dim2<G += Re(conj Z Z2)>": 12.0, "k_dim1_bwd_tma": 40.0}
    it_ms = sum(kt.values())
    achieved = (FWD_BYTES + BWD_BYTES) * ts.px / (it_ms * 1e-3) / 1e9
    traffic, tsrc = None, None
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof):
        try:
            j = json.load(open(prof)).get(name, {})
            traffic, tsrc = j.get("iteration_dram_bytes"), j.get("source")
        except Exception:
            pass
    roofline = {
        "bound": "hbm", "kernel": "one fwd+bwd ADMM iteration = k_dim2t<save> + k_dim1_fwd_tma + k_dim2<accG> + k_dim1_bwd_tma (4 launches)",
        "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
        "traffic": traffic, "traffic_source": tsrc or "not measured in this run (ncu --set full capture under profiles/)",
        "peak_source": pk_src, "algorithmic_bytes_per_plane_pixel_iteration": FWD_BYTES + BWD_BYTES,
        "units_per_launch": ts.px,
        "per_kernel": {k: {"ms": v, "actual_bytes_per_px": actual[k], "actual_GBs": actual[k] * ts.px / (v * 1e-3) / 1e9,
                           "frac_of_peak": actual[k] * ts.px / (v * 1e-3) / 1e9 / pk["hbm_gbs"]} for k, v in kt.items()},
        "other_ms_per_step": f[3] + b[3],
        "timing": "CUDA events around every launch on the launching stream (admmtv_profile_forward / _backward)",
    }
    e2e = e2e_train(c, w, B, steps, warmup, ts)
    extra = {"roofline": roofline, "e2e": e2e, "launches_per_step": ts.launches}
    del ts
    torch.cuda.empty_cache()
    return r, extra


def run_native(args, w, name):
    import torch

    c = setup_dist()
    pk, pk_src = peaks()
    if w["mode"] == "fwd+bwd":
        r, extra = run_train(c, w, name, args.steps, args.warmup, True)
    else:
        r, extra = run_fwd(c, w, name, args.steps, args.warmup), None
    others = {}
    if not args.no_others and name == "cfg2_train":
        for on in OTHERS:
            ow = WORKLOADS[on]
            try:      # >= 3 steps and >= 0.6 s of timed region each (clock samples every 50 ms)
                if ow["mode"] == "fwd+bwd":
                    others[on], _ = run_train(c, ow, on, 3, 3, False, min_ms=600.0)
                else:
                    others[on] = run_fwd(c, ow, on, 3, 3, e2e=(on == "cfg2"), min_ms=600.0)
            except Exception as e:   # an auxiliary shape must never take the headline down
                others[on] = {"error": repr(e)}
            torch.cuda.empty_cache()
    if c.rank == 0:
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": c.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": r["config"],
            "hbm_frac_whole_step": r["frac"], "clocks": r["clocks"],
        }
        if extra:
            e = extra["e2e"]
            px_all = w["B"] * w["P"] * w["N"] * w["M"] * c.world
            line["roofline"] = extra["roofline"]
            e = dict(e)
            f32 = e.pop("f32", None)
            line["e2e"] = {"value": px_all * w["iters"] / (e["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT, **e}
            if f32 is not None:
                if "ms_per_step" in f32:
                    f32 = {"value": px_all * w["iters"] / (f32["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT, **f32}
                line["e2e_f32"] = f32
            line["gpu_launches"] = extra["launches_per_step"] * args.steps
            if c.world > 1:
                line["collective"] = r["collective"]
                line["e2e"]["host_cpus_per_rank"] = getattr(c, "affinity", None)
        elif "e2e" in r:
            line["e2e"] = r["e2e"]
        if c.world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(w)
        if others:
            line["others"] = others
        emit(line)
    if c.world > 1:
        c.dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else this process (or a library under it, e.g. NCCL's
    version banner) writes to file descriptor 1 has been re-routed to stderr by main()."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2_train", choices=sorted(WORKLOADS))
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--graph", type=int, default=None, choices=[0, 1], help="force CUDA-graph replay of the training step off / on")
    args = ap.parse_args()
    global args_graph
    args_graph = args.graph
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_native(args, w, args.workload)


if __name__ == "__main__":
    main()
