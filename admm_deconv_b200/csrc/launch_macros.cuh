// launch_macros.cuh -- kernel launch / dynamic shared memory spellings (nvcc vs test emulation).
#pragma once
#include "compat.cuh"
#ifndef ADMMTV_EMU
#define ADMMTV_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ADMMTV_DYN_SMEM(type, name)                              \
  extern __shared__ __align__(1024) unsigned char name##_raw_[];   \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif
