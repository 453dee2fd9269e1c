// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal CPU emulation of the CUDA execution model, just enough to compile the product's
// kernel sources (admm_deconv_b200/csrc/*.cuh, *.cu) unchanged with g++ and run them with one
// OS thread per CUDA thread (blocks run one after another, __syncthreads is a std::barrier).
// It exists so the kernels' index arithmetic, halo logic and barrier placement can be checked
// against the oracle in the CPU-only test-suite, where no GPU is available.  It is NOT a CPU
// fallback: the product library (libadmmtv.so) is built by nvcc only and never contains or
// loads this code; the emulated library is built into tests/emu/_build/ by tests/emu/build.py.
#pragma once

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };

namespace emu {
struct BlockCtx {
  std::barrier<>* bar;
  unsigned char* dyn_smem;
  // warp shuffle exchange: one slot per thread + one barrier per warp
  std::vector<uint64_t>* shfl;
  std::vector<std::unique_ptr<std::barrier<>>>* wbar;
};
extern thread_local BlockCtx* tls_ctx;
extern thread_local uint3 tls_threadIdx, tls_blockIdx;
extern thread_local dim3 tls_blockDim, tls_gridDim;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace emu

#define threadIdx (emu::tls_threadIdx)
#define blockIdx (emu::tls_blockIdx)
#define blockDim (emu::tls_blockDim)
#define gridDim (emu::tls_gridDim)

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

static inline void __syncthreads() { emu::tls_ctx->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}

template <class T> static inline T __ldg(const T* p) { return *p; }

template <class T>
static inline T emu_shfl_fetch(T v, int src_lane_abs) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  auto* c = emu::tls_ctx;
  unsigned tid = emu::tls_threadIdx.x;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  (*c->shfl)[tid] = bits;
  (*c->wbar)[tid / 32]->arrive_and_wait();
  uint64_t got = (*c->shfl)[src_lane_abs];
  (*c->wbar)[tid / 32]->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  unsigned tid = emu::tls_threadIdx.x;
  return emu_shfl_fetch(v, (int)((tid & ~31u) | ((tid ^ (unsigned)lane_mask) & 31u)));
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int delta) {
  unsigned tid = emu::tls_threadIdx.x;
  unsigned lane = tid & 31u;
  unsigned src = lane + (unsigned)delta < 32u ? lane + (unsigned)delta : lane;
  return emu_shfl_fetch(v, (int)((tid & ~31u) | src));
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src_lane) {
  unsigned tid = emu::tls_threadIdx.x;
  return emu_shfl_fetch(v, (int)((tid & ~31u) | ((unsigned)src_lane & 31u)));
}

static inline float atomicAdd(float* addr, float v) {
  auto* a = reinterpret_cast<std::atomic<float>*>(addr);
  float old = a->load(std::memory_order_relaxed);
  while (!a->compare_exchange_weak(old, old + v)) {}
  return old;
}
static inline double atomicAdd(double* addr, double v) {
  auto* a = reinterpret_cast<std::atomic<double>*>(addr);
  double old = a->load(std::memory_order_relaxed);
  while (!a->compare_exchange_weak(old, old + v)) {}
  return old;
}
static inline int atomicAdd(int* addr, int v) { return __atomic_fetch_add(addr, v, __ATOMIC_RELAXED); }

static inline void sincospi(double x, double* s, double* c) {
  *s = std::sin(M_PI * x);
  *c = std::cos(M_PI * x);
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }

// ---- host runtime subset ------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 3; return cudaSuccess; }
typedef void* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

#define ADMMTV_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define ADMMTV_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::tls_ctx->dyn_smem)
