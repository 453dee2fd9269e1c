"""Builds libadmmtv.so in-tree with nvcc for sm_100a (one translation unit per FFT length so the
8-way parallel build takes about a minute).  ``emulate=True`` is used ONLY by the CPU test-suite
(tests/emu_harness.py): it compiles the same sources with g++ against tests/emu/cuda_emu.h into
tests/emu/_build/ -- that library is never loaded by the product package."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libadmmtv.so")
OBJ = os.path.join(CSRC, "_obj")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_OBJ = os.path.join(EMU_DIR, "_build")
EMU_LIB = os.path.join(EMU_OBJ, "libadmmtv_emu.so")

LOG2_SIZES = tuple(range(5, 13)) + tuple(range(20, 32))  # size ids: 2^5..2^12, then 96,160,...,1920 (fft_core.cuh dim_len)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
GXX_FLAGS = ["-std=c++20", "-O1", "-fPIC", "-pthread", "-DADMMTV_EMU", "-x", "c++"]


def _units(emulate: bool = False):
    """(source, define, object stem)"""
    u = [("admmtv_api.cu", None, "admmtv_api"), ("loss_api.cu", None, "loss_api"), ("inst_generic.cu", None, "inst_generic"),
         ("inst_small.cu", None, "inst_small")]
    if not emulate:   # the host-buffer layer (streams, pinned memory) has no CPU emulation twin
        u.append(("host_api.cu", None, "host_api"))
    for l in LOG2_SIZES:
        u.append(("inst_dim1.cu", l, f"inst_dim1_{l}"))
        u.append(("inst_dim2.cu", l, f"inst_dim2_{l}"))
    return u


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc"))]
    d += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return d


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in deps)


def build(emulate: bool = False, force: bool = False, jobs: int | None = None, verbose: bool = False,
          tag: str | None = None, defines: tuple = (), sizes: tuple | None = None) -> str:
    """tag/defines/sizes: tuning variants (tools/variants.py) -- libadmmtv_<tag>.so built with extra -D flags,
    optionally only for some FFT lengths (other lengths then return ADMMTV_ERR_UNSUPPORTED at link... they are
    compiled as stubs by ADMMTV_STUB)."""
    objdir, lib = (EMU_OBJ, EMU_LIB) if emulate else (OBJ, LIB)
    if tag:
        objdir = os.path.join(CSRC, "_obj_" + tag)
        lib = os.path.join(HERE, f"libadmmtv_{tag}.so")
    os.makedirs(objdir, exist_ok=True)
    # one builder at a time per object directory: concurrent callers (e.g. the ranks of a multi-process test) wait here and
    # then find everything up to date instead of compiling and linking the same files at once
    import fcntl
    with open(os.path.join(objdir, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build_locked(emulate, force, jobs, verbose, sizes, defines, objdir, lib)


def _build_locked(emulate, force, jobs, verbose, sizes, defines, objdir, lib) -> str:
    deps = _deps() + ([os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "cuda_emu.cpp")] if emulate else [])
    jobs = jobs or os.cpu_count() or 4
    logs = {}

    def compile_one(unit):
        src, define, stem = unit
        obj = os.path.join(objdir, stem + ".o")
        if not force and not _stale(obj, deps):
            return obj
        if emulate:
            cmd = ["g++", *GXX_FLAGS, "-I" + EMU_DIR, "-I" + CSRC]
        else:
            cmd = ["nvcc", *NVCC_FLAGS, "-I" + CSRC]
        if define is not None:
            cmd.append(f"-DADMMTV_INST={define}")
            if sizes is not None and define not in sizes:
                cmd.append("-DADMMTV_STUB")
        cmd += [f"-D{x}" for x in defines]
        cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[stem] = r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = list(ex.map(compile_one, _units(emulate)))
    if emulate:
        emu_obj = os.path.join(objdir, "cuda_emu.o")
        if force or _stale(emu_obj, deps):
            subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-pthread", "-I" + EMU_DIR, "-c",
                            os.path.join(EMU_DIR, "cuda_emu.cpp"), "-o", emu_obj], check=True)
        objs.append(emu_obj)
    if force or _stale(lib, objs):
        if emulate:
            cmd = ["g++", "-shared", "-pthread", *objs, "-o", lib]
        else:
            cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed: {' '.join(cmd)}\n{r.stderr}")
    if verbose:
        for k, v in sorted(logs.items()):
            sys.stderr.write(f"--- {k}\n{v}\n")
    if not emulate and logs:
        with open(os.path.join(objdir, "ptxas.log"), "w") as f:
            for k, v in sorted(logs.items()):
                f.write(f"--- {k}\n{v}\n")
    return lib


if __name__ == "__main__":
    print(build(emulate="--emu" in sys.argv, force="--force" in sys.argv, verbose="-v" in sys.argv))
