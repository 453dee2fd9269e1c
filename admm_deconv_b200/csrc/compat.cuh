// compat.cuh -- the one place that knows whether the kernels are being compiled by nvcc for
// sm_100a (the product) or by g++ against tests/emu/cuda_emu.h (a thread-per-CUDA-thread CPU
// emulation used ONLY by the CPU test-suite to exercise the kernels' index logic without a GPU;
// it is never built into, loaded by, or shipped with the product library).
#pragma once

#ifdef ADMMTV_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <stdint.h>

#define ADMMTV_HD __host__ __device__
#define ADMMTV_DI __device__ __forceinline__
