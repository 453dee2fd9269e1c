// tma_prims.cuh -- the TMA / mbarrier spellings (inline PTX) and the host-side tensor-map encoder shared by the dim-1 kernels
// (kernels.cuh: column loads / stores of the spectrum, 128-byte swizzle) and the pipelined dim-2 kernel (kernels_tma.cuh).
#pragma once
#include "compat.cuh"
#ifndef ADMMTV_EMU
#include <cuda.h>

namespace admmtv {

ADMMTV_DI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
ADMMTV_DI void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
ADMMTV_DI void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
ADMMTV_DI void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 3-D tiled TMA load: box at (c0 = float index along dim 1, c1 = column, c2 = pair) -> shared memory, completes on `bar`
ADMMTV_DI void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
ADMMTV_DI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 3-D tiled TMA store: shared memory -> box at (c0, c1, c2); completion through the bulk async-group
ADMMTV_DI void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
ADMMTV_DI void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
ADMMTV_DI void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled tma_encode_fn() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

// Host side: tensor map of `ncols` contiguous columns of M float2 (the [Q][N][M] spectra seen as [Q*N] columns), viewed as
// fp32 [ncols][M/16][32]: box = one whole column as M/16 lines of 128 bytes, written to / read from shared memory with the
// 128-byte swizzle (16-byte chunk index ^= line index mod 8), so that a thread can pull a contiguous 8- or 16-element run of
// a column out of shared memory without bank conflicts (kernels.cuh, dim1_ifft_to_smem_tma).
inline int tma_make_colmap(CUtensorMap* map, const float2* base, int M, size_t ncols) {
  PFN_encodeTiled enc = tma_encode_fn();
  if (!enc || M % 16 != 0 || M / 16 > 256 || ncols == 0 || ncols > 0xffffffffull) return -3;
  const cuuint64_t dims[3] = {32, (cuuint64_t)(M / 16), (cuuint64_t)ncols};
  const cuuint64_t strides[2] = {128, (cuuint64_t)M * sizeof(float2)};
  const cuuint32_t box[3] = {32, (cuuint32_t)(M / 16), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}
// the two column maps of an iteration dim-1 kernel (kernel parameter, __grid_constant__)
struct Dim1Tma {
  CUtensorMap in, out;
};

}  // namespace admmtv
#endif  // !ADMMTV_EMU
