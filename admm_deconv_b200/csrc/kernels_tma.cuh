// kernels_tma.cuh -- k_dim2t: the dim-2 pass (forward FFT -> spectral multiply -> inverse FFT, ops.jl:168 "C .* rfft")
// with TMA tile loads and an mbarrier pipeline.
//
// k_dim2 (kernels.cuh) loads its 16-row x N-column tile with LDG.128 into registers in the first radix pass: the load
// latency of a block can only be hidden by the two other blocks resident on the SM, and every load costs issue slots,
// address arithmetic and registers.  Here one persistent block per SM owns a ring of three tile buffers in shared
// memory (two per compute group).  The TMA unit fills them with 2-D box copies (cp.async.bulk.tensor, SASS UTMALDG; box = TR rows x 256
// columns of the [Q][N][M] float2 spectrum, landing as the [column][row] tile the radix passes use), completion is
// signalled on an mbarrier per buffer, and two independent 256-thread groups (named barriers, no __syncthreads) each
// work through a tile of their own -- all passes in place -- while the third buffer is being filled.  The last inverse
// pass pulls its operands into registers, the buffer is handed straight back to the TMA unit for the tile after next,
// and the results go registers -> global (STG.128) as before.
#pragma once

#include "kernels.cuh"

#ifndef ADMMTV_EMU
#include <cuda.h>

namespace admmtv {

#ifndef ADMMTV_D2T_GROUPS
#define ADMMTV_D2T_GROUPS 3
#endif
#ifndef ADMMTV_D2T_TR9
#define ADMMTV_D2T_TR9 16
#endif
#ifndef ADMMTV_D2T_TR11
#define ADMMTV_D2T_TR11 4
#endif
constexpr int kD2tGroupNT = 256;                // threads per compute group

// Two ring organisations:
//   SHARED : G groups share NB buffers; local tile n lives in buffer n % NB and is worked on by group n % G; the group
//            that finishes tile n hands its buffer to the TMA unit for tile n + NB.  (N = 2048: 64 KB tiles of 4 rows,
//            2 groups, 3 buffers -- the plain kernel needs 8-row tiles there, i.e. ONE 128 KB block per SM with nothing
//            to overlap its loads with, and 4-row tiles make its LDG.128 touch sixteen 32-byte sectors per request.)
//   else   : every group owns two buffers (double buffering): its next tile loads while it works on the current one.
template <int LN>
struct Dim2tCfg {
  static constexpr int N = dim_len(LN);
  static constexpr bool SHARED = LN >= 9;
  static constexpr int TR = LN == 9 ? ADMMTV_D2T_TR9 : (LN == 11 ? ADMMTV_D2T_TR11 : Dim2Cfg<LN>::TR);   // rows per tile
  static constexpr int G = SHARED ? 2 : ADMMTV_D2T_GROUPS;   // independent compute groups per block
  static constexpr int NB = SHARED ? 3 : 2 * G;              // tile buffers
  static constexpr int NT = G * kD2tGroupNT;
  static constexpr int BOXC = N < 256 ? N : 256;          // columns per TMA box (box dimensions are limited to 256)
  static constexpr int NBOX = N / BOXC;
  static constexpr size_t TILE_BYTES = (size_t)N * TR * sizeof(float2);
  static constexpr size_t SMEM = NB * TILE_BYTES + 128;   // + the mbarriers
  static constexpr bool OK = is_pow2(N) && TR >= 2 && SMEM <= 227 * 1024 && ((TR / 2) * (N / plan_radix(N, 0, kP2))) % kD2tGroupNT == 0 &&
                             ((TR / 2) * (N / plan_radix(N, 0, kP2))) / kD2tGroupNT * plan_radix(N, 0, kP2) <= 16;
};

ADMMTV_DI void group_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kD2tGroupNT) : "memory"); }   // named barrier of one group

template <int LN, int S>
ADMMTV_DI void dim2t_fwd_mid(float2* tile, const float2* __restrict__ tw, int lt, int bar) {
  if constexpr (S < plan_stages(dim_len(LN), kP2) - 1) {
    dim2_smem_stage<LN, S, false, kD2tGroupNT, Dim2tCfg<LN>::TR>(tile, tw, lt);
    group_bar(bar);
    dim2t_fwd_mid<LN, S + 1>(tile, tw, lt, bar);
  }
}
template <int LN, int S>
ADMMTV_DI void dim2t_inv_mid(float2* tile, const float2* __restrict__ tw, int lt, int bar) {
  if constexpr (S >= 1) {
    dim2_smem_stage<LN, S, true, kD2tGroupNT, Dim2tCfg<LN>::TR>(tile, tw, lt);
    group_bar(bar);
    dim2t_inv_mid<LN, S - 1>(tile, tw, lt, bar);
  }
}

// variants without gradient accumulation: (MUL, SAVE_Z, FWD_ONLY) as k_dim2
template <int LN, int MUL, bool SAVE_Z, bool FWD_ONLY>
__global__ void __launch_bounds__(Dim2tCfg<LN>::NT, 1) k_dim2t(Dim2Args A, const __grid_constant__ CUtensorMap tmap) {
  using Cfg = Dim2tCfg<LN>;
  constexpr int N = Cfg::N, TR = Cfg::TR, RP = TR / 2, NS = plan_stages(N, kP2), GNT = kD2tGroupNT;
  constexpr int TE = N * TR;   // float2 elements per tile
  using St0 = Stage<N, 0, kP2>;
  constexpr int IT0 = (RP * St0::ITEMS) / GNT;   // stage-0 items per thread
  static_assert((RP * St0::ITEMS) % GNT == 0 && IT0 >= 1 && IT0 * St0::R <= 16, "inverse stage 0 must fit in registers");
  constexpr int G = Cfg::G, NB = Cfg::NB;
  constexpr bool SHARED = Cfg::SHARED;
  ADMMTV_DYN_SMEM(float2, ring);   // [NB][N][TR], then the mbarriers
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)NB * TE);
  const int tid = threadIdx.x, M = A.M;
  const int grp = tid / GNT, lt = tid % GNT, bar = 1 + grp;
  const int row_tiles = M / TR, total = row_tiles * A.Q;
  // this block's tiles: t = blockIdx.x + n * gridDim.x, n = 0 .. nloc-1; group g takes n = g, g + G, ... (its j-th tile).
  // buffer / mbarrier phase of local tile n:  SHARED: n % NB, (n / NB) & 1 ;  else: g*2 + (j & 1), (j >> 1) & 1
  const int nloc = ((int)blockIdx.x < total) ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto buf_of = [&](int n) { return SHARED ? n % NB : (n % G) * 2 + ((n / G) & 1); };
  auto issue = [&](int n) {   // one thread: arm the buffer's mbarrier and start the box copies of local tile n
    if (n >= nloc) return;
    const int t = (int)blockIdx.x + n * (int)gridDim.x;
    const int q = t / row_tiles, i0 = (t % row_tiles) * TR;
    const int b = buf_of(n);
    mbar_expect_tx(full + b, (unsigned)Cfg::TILE_BYTES);
#pragma unroll
    for (int x = 0; x < Cfg::NBOX; ++x)
      tma_load_3d(ring + (size_t)b * TE + (size_t)x * Cfg::BOXC * TR, &tmap, 2 * i0, x * Cfg::BOXC, q, full + b);
  };
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NB; ++b) mbar_init(full + b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    for (int n = 0; n < NB; ++n) issue(n);   // the first NB tiles of the block
  }

  for (int n = grp; n < nloc; n += G) {
    const int t = (int)blockIdx.x + n * (int)gridDim.x;
    const int q = t / row_tiles, i0 = (t % row_tiles) * TR;
    const size_t qoff = (size_t)q * N * M;
    const size_t toff = (size_t)(q / A.Qg) * A.tab_stride;
    const int b = buf_of(n);
    float2* tile = ring + (size_t)b * TE;
    mbar_wait(full + b, (unsigned)((SHARED ? n / NB : n / (2 * G)) & 1));   // the TMA bytes of this tile have landed

    // forward stages 0 .. NS-2 in place
    dim2t_fwd_mid<LN, 0>(tile, A.twN, lt, bar);
    // last forward stage fused with the spectral multiply and the first inverse stage
    dim2_fused_stage<LN, MUL, SAVE_Z, 0, FWD_ONLY, GNT, TR>(tile, nullptr, A, qoff, toff, i0, lt);
    if (!FWD_ONLY) {
      group_bar(bar);
      dim2t_inv_mid<LN, NS - 2>(tile, A.twN, lt, bar);
    }
    // inverse stage 0: every operand of this thread -> registers; then the buffer goes back to the TMA unit
    float4 r[FWD_ONLY ? 1 : IT0][FWD_ONLY ? 1 : St0::R];
    if (!FWD_ONLY) {
#pragma unroll
      for (int u = 0; u < IT0; ++u) {
        const int item = lt + u * GNT, rp = item % RP, wi = item / RP;
#pragma unroll
        for (int m = 0; m < St0::R; ++m) r[u][m] = *reinterpret_cast<const float4*>(tile + (wi + m * St0::STRIDE) * TR + 2 * rp);
      }
    }
    fence_proxy_async();   // this group's generic-proxy accesses to the buffer are ordered before the async-proxy refill
    group_bar(bar);
    if (lt == 0) issue(n + NB);   // the tile that reuses this buffer (SHARED: another group's; else this group's next but one)
    if (!FWD_ONLY) {
      float2* dst = A.out + qoff + i0;
#pragma unroll
      for (int u = 0; u < IT0; ++u) {
        const int item = lt + u * GNT, rp = item % RP, wi = item / RP;
        float2 p[St0::R];
        stage_twiddles<N, 0, true, kP2>(wi, A.twN, p);
        float2 a0[St0::R], a1[St0::R];
#pragma unroll
        for (int m = 0; m < St0::R; ++m) {
          a0[m] = make_float2(r[u][m].x, r[u][m].y);
          a1[m] = make_float2(r[u][m].z, r[u][m].w);
        }
        stage_inv<N, 0, kP2>(a0, p);
        stage_inv<N, 0, kP2>(a1, p);
#pragma unroll
        for (int m = 0; m < St0::R; ++m)
          *reinterpret_cast<float4*>(dst + (size_t)(wi + m * St0::STRIDE) * M + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
      }
    }
  }
}

// Host side: the 3-D tensor map of a [Q][N][M] float2 array, viewed as fp32 [Q][N][2M], box = (2 TR, BOXC, 1).
inline int tma_make_map(CUtensorMap* map, const float2* base, int M, int N, int Q, int TR, int BOXC) {
  PFN_encodeTiled enc = tma_encode_fn();
  if (!enc) return -3;
  const cuuint64_t dims[3] = {(cuuint64_t)2 * M, (cuuint64_t)N, (cuuint64_t)Q};
  const cuuint64_t strides[2] = {(cuuint64_t)M * sizeof(float2), (cuuint64_t)N * M * sizeof(float2)};
  const cuuint32_t box[3] = {(cuuint32_t)2 * TR, (cuuint32_t)BOXC, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

}  // namespace admmtv
#endif  // !ADMMTV_EMU
