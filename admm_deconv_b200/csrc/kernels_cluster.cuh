// kernels_cluster.cuh -- k_dim2c: the dim-2 pass (forward FFT -> spectral multiply -> inverse FFT, ops.jl:168 "C .* rfft")
// of a LONG line split over a thread-block cluster, the shares meeting through distributed shared memory.
//
// k_dim2 (kernels.cuh) needs a whole tile of rows x N columns in one block's shared memory.  For N = 2048 that is 8 rows (128 KB,
// ONE 512-thread block per SM), for N = 4096 only 4 rows: every global access is then a 32-byte run, a quarter of a line.
// Here the line is decimated in time ACROSS a cluster of CS CTAs: CTA c owns the columns j = CS n + c, runs an ordinary
// N/CS-point transform on them in its own 64 KB tile of 16 rows (full 128-byte lines, three CTAs per SM whose load / compute /
// store phases interleave), and the missing radix-CS level is done where the CS sub-spectra Y_c meet (W = exp(-2 pi i / N),
// T = the spectral table, NL = N / CS):
//
//   X[k + m NL]  = sum_c  W_CS^(c m) (W^(c k) Y_c[k])        X'[f] = T[f] X[f]        Y_c'[k] = conj(W^(c k)) sum_m W_CS^(-c m) X'[k + m NL]
//
// Each CTA does this for 1/CS of the positions, reading one operand from every tile of the cluster (ld.shared::cluster) and
// writing all of them back in place, between two cluster barriers; the order of the spectrum inside the kernel is private,
// the tables are read through the column map of the plan they were built for.
//
// Measured on B200 (profiles/r2c_experiments.md): N = 4096 (CS = 8) 293 -> 287 us per pass on 4 x 4096^2: shipped;
// N = 2048: CS = 4 x 16 rows 221 us, CS = 2 x 8 rows 230 us against 216 us for k_dim2<11>: NOT shipped (ADMMTV_D2_CLUSTER_MASK).
// ncu on the 2048 build: DRAM 31 %, issue 41 %, L1 60 %, long-scoreboard and barrier / membar stalls of the two cluster
// barriers: the exchange costs what the finer interleaving gains.
#pragma once

#include "kernels.cuh"

#ifndef ADMMTV_EMU
#include <cooperative_groups.h>

namespace admmtv {

constexpr int kPC = 2;   // plan id of the cluster-local transforms (fft_core.cuh)

#ifndef ADMMTV_D2C_MINB
#define ADMMTV_D2C_MINB 3
#endif
#ifndef ADMMTV_D2C_UNROLL
#define ADMMTV_D2C_UNROLL 2
#endif

#ifndef ADMMTV_D2C_CS
#define ADMMTV_D2C_CS 4    // CTAs per cluster = radix of the level done across distributed shared memory
#endif
#ifndef ADMMTV_D2C_TR
#define ADMMTV_D2C_TR 16   // rows per tile: 16 rows = one full 128-byte line per column and global access
#endif

template <int LN>
struct Dim2cCfg {
  static constexpr int CS = LN >= 12 ? 8 : ADMMTV_D2C_CS;   // N = 4096: eight shares of 512 columns (16 rows x 512 x 8 B = 64 KB each)
  static constexpr int N = dim_len(LN), NL = N / CS;   // line length, length of a CTA's share
  static constexpr int TR = ADMMTV_D2C_TR;             // rows per tile (a multiple of k_dim2's row tile: Dim2Launch::row_tile)
  static constexpr int RP = TR / 2;
  static constexpr int NT = 256;
  static constexpr int NS = plan_stages(NL, kPC);
  static constexpr int RL = plan_radix(NL, NS - 1, kPC);
  static constexpr size_t SMEM = (size_t)NL * TR * sizeof(float2);
  static constexpr bool SWZ = TR < 16;
  static constexpr bool OK = is_pow2(N) && N >= 1024 && (CS == 2 || CS == 4 || CS == 8) && TR >= 2 && TR <= 16 && NS >= 2 &&
                             (RP * (NL / CS)) % NT == 0 && SMEM <= 113 * 1024;
};

// shared-memory column of local position c (Dim2Cfg's XOR swizzle, for the cluster-local plan)
template <int LN>
ADMMTV_DI int d2c_col(int c) {
  using Cfg = Dim2cCfg<LN>;
  if constexpr (!Cfg::SWZ) return c;
  else {
    constexpr int SH = Cfg::RL == 32 ? 5 : (Cfg::RL == 16 ? 4 : (Cfg::RL == 8 ? 3 : 2));
    return c ^ ((c >> SH) & (16 / Cfg::TR - 1));
  }
}

// twiddle powers of work item wi of stage S of the LOCAL transform: W_NL^t = W_N^(CS t)
template <int LN, int S, bool INV>
ADMMTV_DI void d2c_twiddles(int wi, const float2* __restrict__ twN, float2* p) {
  using St = Stage<Dim2cCfg<LN>::NL, S, kPC>;
  if constexpr (St::HAS_TW) {
    float2 w = twN[Dim2cCfg<LN>::CS * St::tindex(wi)];
    if (INV) w.y = -w.y;
    twiddle_powers<St::R>(w, p);
  }
}

template <int LN, int S, bool INV>
ADMMTV_DI void d2c_smem_stage(float2* tile, const float2* __restrict__ twN, int tid) {
  using Cfg = Dim2cCfg<LN>;
  constexpr int NL = Cfg::NL, TR = Cfg::TR, RP = Cfg::RP;
  using St = Stage<NL, S, kPC>;
  for (int item = tid; item < RP * St::ITEMS; item += Cfg::NT) {
    const int rp = item % RP, wi = item / RP;
    float2 p[St::R];
    d2c_twiddles<LN, S, INV>(wi, twN, p);
    float2 a0[St::R], a1[St::R];
    const int base = St::base(wi);
#pragma unroll
    for (int m = 0; m < St::R; ++m) {
      const float4 v = *reinterpret_cast<const float4*>(tile + d2c_col<LN>(base + m * St::STRIDE) * TR + 2 * rp);
      a0[m] = make_float2(v.x, v.y);
      a1[m] = make_float2(v.z, v.w);
    }
    if (INV) {
      stage_inv<NL, S, kPC>(a0, p);
      stage_inv<NL, S, kPC>(a1, p);
    } else {
      stage_fwd<NL, S, kPC>(a0, p);
      stage_fwd<NL, S, kPC>(a1, p);
    }
#pragma unroll
    for (int m = 0; m < St::R; ++m)
      *reinterpret_cast<float4*>(tile + d2c_col<LN>(base + m * St::STRIDE) * TR + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
  }
}
template <int LN, int S>
ADMMTV_DI void d2c_fwd_rest(float2* tile, const float2* __restrict__ twN, int tid) {   // stages S .. NS-1, a block barrier between them
  if constexpr (S < Dim2cCfg<LN>::NS) {
    d2c_smem_stage<LN, S, false>(tile, twN, tid);
    if constexpr (S + 1 < Dim2cCfg<LN>::NS) __syncthreads();
    d2c_fwd_rest<LN, S + 1>(tile, twN, tid);
  }
}
template <int LN, int S>
ADMMTV_DI void d2c_inv_rest(float2* tile, const float2* __restrict__ twN, int tid) {   // stages S .. 1, a block barrier after each
  if constexpr (S >= 1) {
    d2c_smem_stage<LN, S, true>(tile, twN, tid);
    __syncthreads();
    d2c_inv_rest<LN, S - 1>(tile, twN, tid);
  }
}

// MUL as k_dim2: 0 = real table ctab, 1 = complex table ktab, 2 = its conjugate (the dispatch ships 0 and 1, the variants the GPU suite exercises at N = 4096)
template <int LN, int MUL>
__global__ void __cluster_dims__(Dim2cCfg<LN>::CS, 1, 1) __launch_bounds__(Dim2cCfg<LN>::NT, ADMMTV_D2C_MINB) k_dim2c(Dim2Args A) {
  namespace cg = cooperative_groups;
  using Cfg = Dim2cCfg<LN>;
  constexpr int N = Cfg::N, NL = Cfg::NL, TR = Cfg::TR, RP = Cfg::RP, NT = Cfg::NT, CS = Cfg::CS;
  using St0 = Stage<NL, 0, kPC>;
  ADMMTV_DYN_SMEM(float2, tile);   // [NL][TR]: this CTA's share of the line (columns CS n + c), all passes in place
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  float2* tl[CS];                  // tl[r] = the tile of cluster rank r (sub-spectrum Y_r)
#pragma unroll
  for (int r = 0; r < CS; ++r) tl[r] = cluster.map_shared_rank(tile, r);
  const int tid = threadIdx.x, M = A.M;
  const int i0 = (int)(blockIdx.x / CS) * TR;

  for (int q = blockIdx.y; q < A.Q; q += gridDim.y) {
    const size_t qoff = (size_t)q * N * M;
    const size_t toff = (size_t)(q / A.Qg) * A.tab_stride;
    const float2* src = A.in + qoff + i0 + (size_t)c * M;   // column CS n + c  ->  src + n * CS * M
    // forward stage 0: global -> registers -> shared
    for (int item = tid; item < RP * St0::ITEMS; item += NT) {
      const int rp = item % RP, wi = item / RP;
      float2 p[St0::R];
      d2c_twiddles<LN, 0, false>(wi, A.twN, p);
      float2 a0[St0::R], a1[St0::R];
#pragma unroll
      for (int m = 0; m < St0::R; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)(wi + m * St0::STRIDE) * CS * M + 2 * rp);
        a0[m] = make_float2(v.x, v.y);
        a1[m] = make_float2(v.z, v.w);
      }
      stage_fwd<NL, 0, kPC>(a0, p);
      stage_fwd<NL, 0, kPC>(a1, p);
#pragma unroll
      for (int m = 0; m < St0::R; ++m)
        *reinterpret_cast<float4*>(tile + d2c_col<LN>(wi + m * St0::STRIDE) * TR + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
    }
    __syncthreads();
    d2c_fwd_rest<LN, 1>(tile, A.twN, tid);
    cluster.sync();   // every sub-spectrum complete and visible to the cluster

    // the radix-CS level across the cluster, fused with the spectral multiply and its own transpose.
    // CTA c takes the local positions [c NL/CS, (c+1) NL/CS): one operand from every tile, all written back in place.
    constexpr int UNR = ADMMTV_D2C_UNROLL;
#pragma unroll UNR
    for (int item = tid; item < RP * (NL / CS); item += NT) {
      const int rp = item % RP, pl = c * (NL / CS) + item / RP;
      const int k = pos_to_freq(NL, pl, true, kPC);              // frequency (mod NL) held at this position of every tile
      const int so = d2c_col<LN>(pl) * TR + 2 * rp;
      float2 va[CS], vb[CS], w[CS];                               // rows 2 rp and 2 rp + 1
#pragma unroll
      for (int r = 0; r < CS; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(tl[r] + so);
        va[r] = make_float2(v.x, v.y);
        vb[r] = make_float2(v.z, v.w);
      }
      twiddle_powers<CS>(A.twN[k], w);                            // W_N^(r k)
#pragma unroll
      for (int r = 1; r < CS; ++r) {
        va[r] = cmul(va[r], w[r]);
        vb[r] = cmul(vb[r], w[r]);
      }
      Dft<CS, false>::run(va);                                    // va[m] = X[k + m NL]
      Dft<CS, false>::run(vb);
#pragma unroll
      for (int m = 0; m < CS; ++m) {
        const size_t g = toff + (size_t)freq_to_pos(N, k + m * NL, kP2) * M + i0 + 2 * rp;   // table column of that frequency
        if (MUL == 0) {
          const float2 cc = *reinterpret_cast<const float2*>(A.ctab + g);
          va[m] = cscale(va[m], cc.x);
          vb[m] = cscale(vb[m], cc.y);
        } else {
          const float sgn = MUL == 2 ? -1.f : 1.f;
          const float4 kk = *reinterpret_cast<const float4*>(A.ktab + g);
          va[m] = cmul(va[m], make_float2(kk.x, sgn * kk.y));
          vb[m] = cmul(vb[m], make_float2(kk.z, sgn * kk.w));
        }
      }
      Dft<CS, true>::run(va);
      Dft<CS, true>::run(vb);
#pragma unroll
      for (int r = 0; r < CS; ++r) {
        if (r > 0) {
          const float2 wc = make_float2(w[r].x, -w[r].y);
          va[r] = cmul(va[r], wc);
          vb[r] = cmul(vb[r], wc);
        }
        *reinterpret_cast<float4*>(tl[r] + so) = make_float4(va[r].x, va[r].y, vb[r].x, vb[r].y);
      }
    }
    cluster.sync();   // the partner's writes into this tile have landed; no remote access after this point

    d2c_inv_rest<LN, Cfg::NS - 1>(tile, A.twN, tid);
    // inverse stage 0: shared -> registers -> global
    float2* dst = A.out + qoff + i0 + (size_t)c * M;   // column CS n + c
    for (int item = tid; item < RP * St0::ITEMS; item += NT) {
      const int rp = item % RP, wi = item / RP;
      float2 p[St0::R];
      d2c_twiddles<LN, 0, true>(wi, A.twN, p);
      float2 a0[St0::R], a1[St0::R];
#pragma unroll
      for (int m = 0; m < St0::R; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(tile + d2c_col<LN>(wi + m * St0::STRIDE) * TR + 2 * rp);
        a0[m] = make_float2(v.x, v.y);
        a1[m] = make_float2(v.z, v.w);
      }
      stage_inv<NL, 0, kPC>(a0, p);
      stage_inv<NL, 0, kPC>(a1, p);
#pragma unroll
      for (int m = 0; m < St0::R; ++m)
        *reinterpret_cast<float4*>(dst + (size_t)(wi + m * St0::STRIDE) * CS * M + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
    }
    __syncthreads();   // the tile is rewritten by the next pair's first pass
  }
}

}  // namespace admmtv
#endif  // !ADMMTV_EMU
