// kernels.cuh -- the hand-written sm_100a kernels of the ADMM-TV path.
//
// Data layout in HBM (DESIGN.md "Layout"): the S = P*B independent image planes are processed
// as Q = ceil(S/2) *plane pairs*: planes (2q, 2q+1) are the real and imaginary part of one
// complex image.  The x-update operator A = F^-1 C F is real, hence C-linear, so one full
// complex 2-D FFT of the pair transforms both planes at once with no Hermitian bookkeeping.
// Every internal array is float2 [Q][N][M] (dim 1 = M contiguous):
//   bpk   = H^T y                               (pair-packed, spatial)
//   spec* = spectra between the two FFT passes   (rows = dim-1 frequencies, digit-reversed)
//   v     = D x + u, the ONLY iteration state:   z = shrink(v), u = v - z   [Q][2][N][M]
//
// One ADMM iteration (ops.jl:168-173) is two launches:
//   k_dim2     : dim-2 FFT -> x C (spectral division, ops.jl:168) -> dim-2 IFFT, per 16-row tile
//   k_dim1_fwd : dim-1 IFFT -> x ; D(x) (ops.jl:169) ; z/u update (ops.jl:171,173) ;
//                r = H^T y + rho D^T(z-u) (ops.jl:168 of the NEXT iteration) ; dim-1 FFT
#pragma once

#include "fft_core.cuh"
#include "args.cuh"
#include "launch_macros.cuh"
#include "reduce.cuh"
#include "tma_prims.cuh"

#ifndef ADMMTV_UNROLL_ITEMS
#define ADMMTV_UNROLL_ITEMS 1
#endif

namespace admmtv {

constexpr int kUnrollItems = ADMMTV_UNROLL_ITEMS;

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
template <int R>
struct TwP {
  float2 p[R];
};

// Distribute (work item wi of a line, line c) pairs over the block so that a thread keeps the
// same wi (hence the same twiddles) whenever the block is at least one line wide.
template <int ITEMS, int NT, class Prep, class Body>
ADMMTV_DI void for_items(int tid, int nlines, Prep prep, Body body) {
  if constexpr (NT % ITEMS != 0 && ITEMS % NT != 0) {
    // generic mapping (non-power-of-two lengths): flat (line, item) index, twiddles per item
    for (int it = tid; it < nlines * ITEMS; it += NT) {
      const int wi = it % ITEMS, c = it / ITEMS;
      const auto ctx = prep(wi);
      body(wi, c, ctx);
    }
  } else if constexpr (NT >= ITEMS) {
    const int wi = tid % ITEMS;
    const auto ctx = prep(wi);
#pragma unroll(kUnrollItems)
    for (int c = tid / ITEMS; c < nlines; c += NT / ITEMS) body(wi, c, ctx);
  } else {
    for (int wi = tid; wi < ITEMS; wi += NT) {
      const auto ctx = prep(wi);
      for (int c = 0; c < nlines; ++c) body(wi, c, ctx);
    }
  }
}

template <int R>
ADMMTV_DI void load_contig(const float2* __restrict__ src, float2* a) {
  static_assert(R % 2 == 0, "even radix");
#pragma unroll
  for (int m = 0; m < R / 2; ++m) {
    const float4 v = *reinterpret_cast<const float4*>(src + 2 * m);
    a[2 * m] = make_float2(v.x, v.y);
    a[2 * m + 1] = make_float2(v.z, v.w);
  }
}
template <int R>
ADMMTV_DI void store_contig(float2* __restrict__ dst, const float2* a) {
#pragma unroll
  for (int m = 0; m < R / 2; ++m)
    *reinterpret_cast<float4*>(dst + 2 * m) = make_float4(a[2 * m].x, a[2 * m].y, a[2 * m + 1].x, a[2 * m + 1].y);
}

ADMMTV_DI float act_apply(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return fminf(fmaxf(v, 0.f), 6.f);
  if (act == 3) return fminf(fmaxf(v, 0.f), 1.f);
  return v;
}
// derivative of the activation expressed through its OUTPUT (relu family: 1 strictly inside)
ADMMTV_DI float act_grad_from_out(float o, int act) {
  if (act == 1) return o > 0.f ? 1.f : 0.f;
  if (act == 2) return (o > 0.f && o < 6.f) ? 1.f : 0.f;
  if (act == 3) return (o > 0.f && o < 1.f) ? 1.f : 0.f;
  return 1.f;
}

// ------------------------------------------------------------------------------------------
// dim-1 (contiguous) FFT passes over a tile of columns held in shared memory
// ------------------------------------------------------------------------------------------
#ifndef ADMMTV_TC9
#define ADMMTV_TC9 10
#endif
#ifndef ADMMTV_PREFETCH
#define ADMMTV_PREFETCH 1
#endif
#ifndef ADMMTV_PF_NEXT
#define ADMMTV_PF_NEXT 0  // L2 prefetch for the block that takes over the SM slot next (distance in blocks); measured SLOWER (cfg2: dim2 99 -> 136 us at 444), kept off
#endif
#ifndef ADMMTV_PF_NEXT2_11
#define ADMMTV_PF_NEXT2_11 0  // the same for the dim-2 kernel at N = 2048 (one 128 KB block per SM: nothing else hides its loads)
#endif
#ifndef ADMMTV_PRELOAD1
#define ADMMTV_PRELOAD1 1
#endif
#ifndef ADMMTV_NT9
#define ADMMTV_NT9 256
#endif
#ifndef ADMMTV_MINB9
#define ADMMTV_MINB9 3
#endif
#ifndef ADMMTV_CHUNK9
#define ADMMTV_CHUNK9 2
#endif
#ifndef ADMMTV_NT11
#define ADMMTV_NT11 512
#endif
#ifndef ADMMTV_TC11
#define ADMMTV_TC11 6
#endif
#ifndef ADMMTV_TR11
#define ADMMTV_TR11 8
#endif
#ifndef ADMMTV_CHUNK8
#define ADMMTV_CHUNK8 8
#endif
#ifndef ADMMTV_TC8
#define ADMMTV_TC8 10
#endif
#ifndef ADMMTV_TR9
#define ADMMTV_TR9 16
#endif
#ifndef ADMMTV_TC8B
#define ADMMTV_TC8B 18   // backward tile at M = 256: 16 + 2 columns (12 % halo instead of 25 %): k_dim1_bwd_tma<8> 392 -> 365 us on 256 x 256^2 x 3
#endif
#ifndef ADMMTV_TC7
#define ADMMTV_TC7 10   // 128-row planes: 8 + 2 columns, chunk 4, 8 blocks/SM: dim-1 111 -> 94 us on 1024 x 128^2 (gpurun_out/v4_128b.log)
#endif
#ifndef ADMMTV_CHUNK7
#define ADMMTV_CHUNK7 4
#endif
#ifndef ADMMTV_MINB7
#define ADMMTV_MINB7 8
#endif
#ifndef ADMMTV_TR7
#define ADMMTV_TR7 16
#endif
#ifndef ADMMTV_MINB8
#define ADMMTV_MINB8 3   // 256-row planes: 3 blocks/SM (<= 85 registers): dim-1 48.0 -> 46.0 us on 96 planes of 256^2 (gpurun_out/v6_mb8.log)
#endif
#ifndef ADMMTV_TR8
#define ADMMTV_TR8 16
#endif
#ifndef ADMMTV_NT2
#define ADMMTV_NT2 256
#endif
#ifndef ADMMTV_NT2_MAX
#define ADMMTV_NT2_MAX 512
#endif
#ifndef ADMMTV_MINB2
#define ADMMTV_MINB2 1
#endif
#ifndef ADMMTV_D2_CPRE
#define ADMMTV_D2_CPRE 0   // hoist the C-table loads of the fused dim-2 stage above its butterflies
#endif
#ifndef ADMMTV_MINB2_11
#define ADMMTV_MINB2_11 1
#endif
#ifndef ADMMTV_D2_SWZ
#define ADMMTV_D2_SWZ 1   // XOR column swizzle of dim-2 tiles narrower than 16 rows (Dim2Cfg::SWZ)
#endif
#ifndef ADMMTV_REV1
#define ADMMTV_REV1 1   // iteration dim-1 kernels walk (pair, tile) in DESCENDING order: they start on the spectra the dim-2 kernel wrote last (still in L2),
                        // and the ascending dim-2 kernel that follows starts on what they wrote last: cfg2 iteration 342 -> 336 us, 256^2 348 -> 344, 2048^2 604 -> 599
#endif
// block index of the iteration dim-1 kernels (k_dim1_fwd / k_dim1_bwd); the dim-2 kernels always ascend
ADMMTV_DI int dim1_bid() {
#if ADMMTV_REV1
  return (int)(gridDim.x - 1 - blockIdx.x);
#else
  return (int)blockIdx.x;
#endif
}
// Asynchronous bulk prefetch of a contiguous global range into L2 (TMA unit, no registers, no
// shared memory): issued for the stencil phase's inputs while the FFT passes run.
ADMMTV_DI void l2_prefetch_bulk(const void* p, unsigned bytes) {
#if !defined(ADMMTV_EMU) && ADMMTV_PREFETCH
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
  (void)p;
  (void)bytes;
#endif
}

// block size of the dim-1 kernels for a non-power-of-two length: M / RPT threads, a multiple of 32, at most 320
ADMMTV_HD constexpr int dim1_nt_generic(int M) {
  for (int r = 1; r <= 16; ++r)
    if (M % r == 0 && (M / r) % 32 == 0 && M / r <= 320) return M / r;
  return 32;
}
ADMMTV_HD constexpr int dim1_tc_generic(int M) {  // tile columns (incl. 2 halo) within ~64 KB of shared memory
  return M <= 192 ? 34 : (M <= 384 ? 18 : (M <= 768 ? 10 : 6));
}

template <int LM>
struct Dim1Cfg {
  static constexpr int M = dim_len(LM);
  static constexpr bool POW2 = is_pow2(M);
  static constexpr int NT = !POW2 ? dim1_nt_generic(M)
                                  : (LM <= 8 ? M : (LM == 9 ? ADMMTV_NT9 : (LM == 10 ? 256 : (LM == 11 ? ADMMTV_NT11 : 512))));
  static constexpr int MINB = LM == 9 ? ADMMTV_MINB9 : (LM == 7 ? ADMMTV_MINB7 : (LM == 8 ? ADMMTV_MINB8 : 1));
  static constexpr int RPT = M / NT;                  // rows per thread in the stencil sweep
  static constexpr int CHUNK = LM == 9 ? ADMMTV_CHUNK9 : (LM == 8 ? ADMMTV_CHUNK8 : (LM == 7 ? ADMMTV_CHUNK7 : (RPT >= 8 ? 1 : 8 / RPT)));  // columns between barriers
  static constexpr int CHUNKB = RPT >= 4 ? 1 : 4 / RPT;  // backward sweep: RPT * CHUNKB <= 4 keeps its hoisted loads in registers
  // tile columns including the 2 halo columns
  static constexpr int TC = !POW2 ? dim1_tc_generic(M)
                                  : (LM <= 6 ? 34 : (LM == 7 ? ADMMTV_TC7 : (LM == 8 ? ADMMTV_TC8 : (LM == 9 ? ADMMTV_TC9 : (LM == 11 ? ADMMTV_TC11 : 6)))));
  static constexpr int CO = TC - 2;                   // output columns per block
  static constexpr size_t SMEM = (size_t)TC * M * sizeof(float2);
  // the backward iteration kernel (k_dim1_bwd, 2 blocks of 128 registers per SM at M = 256) may use a wider tile
  static constexpr int TCB = LM == 8 ? ADMMTV_TC8B : TC;
  static constexpr int COB = TCB - 2;
  static constexpr size_t SMEMB = (size_t)TCB * M * sizeof(float2);
  static_assert(M % NT == 0 && NT % 32 == 0, "dim-1 block must tile the column in whole warps");
};

// k_dim1_fwd, anisotropic (MODE 0): block size and row blocks per thread.  At M = 2048 the tile (6 columns = 96 KB) and the
// 118-128 registers of the sweep allow ONE 512-thread block per SM, whose IFFT-load / stencil / FFT-store phases then run
// back to back with nothing to overlap them (ncu: 24 % warps active, DRAM 45 %, issue 37 %).  With RB = 2 the block has half
// the threads; every thread sweeps rows [tid*RPT, +RPT) and then [M/2 + tid*RPT, +RPT): same registers per thread, two
// blocks per SM out of phase with each other, and the FFT passes (2 columns per round) leave no thread idle.
#ifndef ADMMTV_RB11
#define ADMMTV_RB11 2
#endif
template <int LM, int MODE>
struct Dim1FwdCfg {
  using C = Dim1Cfg<LM>;
  static constexpr int RB = (LM == 11 && MODE == 0) ? ADMMTV_RB11 : 1;
  static constexpr int NT = C::NT / RB;
  static constexpr int MINB = RB > 1 ? RB : C::MINB;
  static_assert(RB == 1 || RB == 2, "row blocks: 1 or 2");
};

// row index modulo M (circular boundary of the difference operators)
template <int M>
ADMMTV_DI int wrapm(int i) {
  if constexpr (is_pow2(M)) return i & (M - 1);
  else return i < 0 ? i + M : (i >= M ? i - M : i);
}

// Shared-memory placement of (column c, row i): X[c*M + (i ^ swz_row(i) ^ swz_col(c))].
// The XOR swizzle folds higher index bits into the 4 bits that select the 8-byte bank pair, so
// that every FFT pass of the plan (strides M/R, M/R^2, ..., 1) is bank-conflict free; the maps
// below were found by tools/bank_model.py.  swz is GF(2)-linear and only moves bits downwards, so
// it is a bijection, and for an element base|off with disjoint bit fields the placement splits as
// place(base) ^ place(off) -- the per-butterfly offsets m*STRIDE fold into compile-time XOR masks.
#ifndef ADMMTV_SWZ
#define ADMMTV_SWZ 1
#endif
template <int LM>
ADMMTV_HD constexpr int swz_row(int i) {
#if ADMMTV_SWZ
  if (LM >= 20) return 0;   // non-power-of-two lengths: plain layout
  if (LM == 5) return ((i >> 2) & 15) ^ ((i >> 3) & 15);
  if (LM == 6 || LM == 7) return (i >> 3) & 15;
  if (LM == 8) return (i >> 4) & 15;
  if (LM == 9 || LM == 10) return ((i >> 2) & 15) ^ ((i >> 3) & 15);
  return ((i >> 4) & 15) ^ ((i >> 5) & 8);
#else
  return 0;
#endif
}
template <int LM>
ADMMTV_HD constexpr int swz_col(int c) {
#if ADMMTV_SWZ
  if (LM >= 20) return 0;
  if (LM == 5) return (c << 2) & 15;
  if (LM == 6 || LM == 7) return (c << 3) & 15;
#endif
  return 0;
}
template <int LM>
ADMMTV_DI int sidx(int c, int i) {
  return c * dim_len(LM) + (i ^ swz_row<LM>(i) ^ swz_col<LM>(c));
}
// placement of a compile-time offset whose bits are disjoint from the base it is combined with
template <int LM>
ADMMTV_HD constexpr int soff(int off) {
  return off ^ swz_row<LM>(off);
}
// element (base + off) of a butterfly given the placed base: XOR-combine under the swizzle (power-of-two
// lengths, disjoint bit fields), plain addition otherwise
template <int LM>
ADMMTV_DI int scomb(int pb, int off) {
  if constexpr (is_pow2(dim_len(LM))) return pb ^ (off ^ swz_row<LM>(off));
  else return pb + off;
}

template <int LM, int NT, int S, bool INV>
ADMMTV_DI void dim1_smem_stage(float2* X, int ncols, const float2* __restrict__ tw, int tid) {
  constexpr int M = dim_len(LM);
  using St = Stage<M, S>;
  for_items<St::ITEMS, NT>(
      tid, ncols,
      [&](int wi) {
        TwP<St::R> t;
        stage_twiddles<M, S, INV>(wi, tw, t.p);
        return t;
      },
      [&](int wi, int c, const TwP<St::R>& t) {
        float2 a[St::R];
        const int pb = sidx<LM>(c, St::base(wi));
#pragma unroll
        for (int m = 0; m < St::R; ++m) a[m] = X[scomb<LM>(pb, m * St::STRIDE)];
        if (INV) stage_inv<M, S>(a, t.p);
        else stage_fwd<M, S>(a, t.p);
#pragma unroll
        for (int m = 0; m < St::R; ++m) X[scomb<LM>(pb, m * St::STRIDE)] = a[m];
      });
}

template <int LM, int NT, int S>
ADMMTV_DI void dim1_inv_stages_down(float2* X, int ncols, const float2* __restrict__ tw, int tid) {
  if constexpr (S >= 0) {
    dim1_smem_stage<LM, NT, S, true>(X, ncols, tw, tid);
    __syncthreads();
    dim1_inv_stages_down<LM, NT, S - 1>(X, ncols, tw, tid);
  }
}
template <int LM, int NT, int S>
ADMMTV_DI void dim1_fwd_stages_up(float2* X, int ncols, const float2* __restrict__ tw, int tid) {
  if constexpr (S < plan_stages(dim_len(LM)) - 1) {
    dim1_smem_stage<LM, NT, S, false>(X, ncols, tw, tid);
    __syncthreads();
    dim1_fwd_stages_up<LM, NT, S + 1>(X, ncols, tw, tid);
  }
}

// Inverse dim-1 FFT of `ncols` spectrum columns (digit-reversed rows) -> natural-order columns
// in X.  The first pass reads global memory directly (contiguous R-element runs).  Ends synced.
template <int LM, int NT, class ColPtr>
ADMMTV_DI void dim1_ifft_to_smem(float2* X, int ncols, ColPtr colptr, const float2* __restrict__ tw, int tid) {
  constexpr int M = dim_len(LM), NS = plan_stages(M);
  using St = Stage<M, NS - 1>;
  static_assert(St::STRIDE == 1, "last plan stage must be contiguous");
  constexpr int TCMAX = Dim1Cfg<LM>::TC > Dim1Cfg<LM>::TCB ? Dim1Cfg<LM>::TC : Dim1Cfg<LM>::TCB;   // widest tile any caller passes
  if constexpr (ADMMTV_PRELOAD1 && NT >= St::ITEMS && (TCMAX + NT / St::ITEMS - 1) / (NT / St::ITEMS) * St::R <= 24) {
    // all of this thread's global loads of the pass are issued before the first butterfly
    constexpr int CSTEP = NT / St::ITEMS, ROUNDS = (TCMAX + CSTEP - 1) / CSTEP;
    const int wi = tid % St::ITEMS, c0 = tid / St::ITEMS;
    float2 a[ROUNDS][St::R];
#pragma unroll
    for (int n = 0; n < ROUNDS; ++n)
      if (c0 + n * CSTEP < ncols) load_contig<St::R>(colptr(c0 + n * CSTEP) + wi * St::R, a[n]);
#pragma unroll
    for (int n = 0; n < ROUNDS; ++n) {
      const int c = c0 + n * CSTEP;
      if (c < ncols) {
        Dft<St::R, true>::run(a[n]);
        const int pb = sidx<LM>(c, wi * St::R);
#pragma unroll
        for (int m = 0; m < St::R; ++m) X[scomb<LM>(pb, m)] = a[n][m];
      }
    }
  } else {
    for_items<St::ITEMS, NT>(
        tid, ncols, [&](int) { return 0; },
        [&](int wi, int c, int) {
          float2 a[St::R];
          load_contig<St::R>(colptr(c) + wi * St::R, a);
          Dft<St::R, true>::run(a);
          const int pb = sidx<LM>(c, wi * St::R);
#pragma unroll
          for (int m = 0; m < St::R; ++m) X[scomb<LM>(pb, m)] = a[m];
        });
  }
  __syncthreads();
  dim1_inv_stages_down<LM, NT, NS - 2>(X, ncols, tw, tid);
}

// Forward dim-1 FFT of `ncols` natural-order columns in X (must be synced) -> global spectrum
// columns (digit-reversed rows).  The last pass writes global memory directly.
template <int LM, int NT, class ColPtr>
ADMMTV_DI void dim1_fft_from_smem(float2* X, int ncols, ColPtr colptr, const float2* __restrict__ tw, int tid) {
  constexpr int M = dim_len(LM), NS = plan_stages(M);
  dim1_fwd_stages_up<LM, NT, 0>(X, ncols, tw, tid);
  using St = Stage<M, NS - 1>;
  for_items<St::ITEMS, NT>(
      tid, ncols, [&](int) { return 0; },
      [&](int wi, int c, int) {
        float2 a[St::R];
        const int pb = sidx<LM>(c, wi * St::R);
#pragma unroll
        for (int m = 0; m < St::R; ++m) a[m] = X[scomb<LM>(pb, m)];
        Dft<St::R, false>::run(a);
        store_contig<St::R>(colptr(c) + wi * St::R, a);
      });
}

// ------------------------------------------------------------------------------------------
// TMA column loads / stores for the dim-1 transforms (north_star: "FFT staged in shared memory, TMA tile loads").
// The stride-1 radix pass of a dim-1 transform gives every thread a contiguous run of R = 8 or 16 elements of a column.
// Straight from / to global memory (dim1_ifft_to_smem / dim1_fft_from_smem above) that is R/2 LDG.128 / STG.128 per
// thread whose lanes sit 64 or 128 bytes apart: every warp request touches 16 or 32 different 128-byte lines, and ncu
// shows these two passes producing about half of the L1 data-pipe wavefronts of the iteration kernels, which run at
// 62-72 % of that pipe (their tightest resource; profiles/r2b_*).  Here the TMA unit moves whole columns
// (cp.async.bulk.tensor, one box = M/16 lines of 128 bytes, 128-byte swizzle) between global and shared memory, and the
// radix pass reads / writes its runs in shared memory: with the swizzle (16-byte chunk ^= line mod 8) the 8 lanes of a
// quarter-warp hit 8 different chunk slots, i.e. 4 wavefronts per 512-byte request, the minimum.  Both layouts (the TMA
// swizzle and the pass-conflict-free XOR placement sidx) permute elements inside one 128-byte line, so the pass runs in place;
// the lanes that share a line (two for R = 8) are neighbours in a warp and separated by __syncwarp.
// ------------------------------------------------------------------------------------------
#ifdef ADMMTV_EMU
struct Dim1Tma {};
#else
#ifndef ADMMTV_D1_TMA
#define ADMMTV_D1_TMA 1
#endif
template <int LM>
constexpr bool kDim1TmaOk = ADMMTV_D1_TMA && LM >= 7 && LM <= 12;   // 128 .. 4096: stride-1 pass of radix 8 or 16, M/16 <= 256 lines

// byte offset (from the tile base) of logical 16-byte chunk g of the 128-byte line at byte offset lineoff
ADMMTV_DI unsigned tma_chunk_off(unsigned xbase, unsigned lineoff, int g) {
  return lineoff + ((unsigned)(g ^ (int)(((xbase + lineoff) >> 7) & 7u)) << 4);
}
// (work item, column) pairs of the stride-1 pass with a warp-uniform trip count: body(wi, c, active)
template <int ITEMS, int NT, class Body>
ADMMTV_DI void for_items_uniform(int tid, int ncols, Body body) {
  if constexpr (NT >= ITEMS) {
    constexpr int CSTEP = NT / ITEMS;
    const int wi = tid % ITEMS, c0 = tid / ITEMS;
    for (int cb = 0; cb < ncols; cb += CSTEP) body(wi, cb + c0, cb + c0 < ncols);
  } else {
    static_assert(ITEMS % NT == 0, "items tile the block");
    for (int wi = tid; wi < ITEMS; wi += NT)
      for (int c = 0; c < ncols; ++c) body(wi, c, true);
  }
}

// Inverse dim-1 FFT of `ncols` spectrum columns -> natural-order columns in X; colidx(c) = index of tile column c among
// the columns of the map.  Ends synced.
template <int LM, int NT, class ColIdx>
ADMMTV_DI void dim1_ifft_to_smem_tma(float2* X, int ncols, const CUtensorMap* map, ColIdx colidx, unsigned long long* bar,
                                     const float2* __restrict__ tw, int tid) {
  constexpr int M = dim_len(LM), NS = plan_stages(M);
  using St = Stage<M, NS - 1>;
  constexpr int R = St::R, CPI = R / 2;   // 16-byte chunks per work item
  static_assert(St::STRIDE == 1 && (R == 8 || R == 16), "stride-1 pass of radix 8 or 16");
  const unsigned xbase = smem_u32(X);
  if (tid == 0) {
    if (xbase & 127u) __trap();   // TMA destination alignment
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, (unsigned)(ncols * M * sizeof(float2)));
    for (int c = 0; c < ncols; ++c) tma_load_3d(X + (size_t)c * M, map, 0, 0, colidx(c), bar);
  }
  __syncthreads();   // the barrier is initialised before anyone waits on it
  mbar_wait(bar, 0);
  char* Xb = reinterpret_cast<char*>(X);
  for_items_uniform<St::ITEMS, NT>(tid, ncols, [&](int wi, int c, bool act) {
    float2 a[R];
    if (act) {
      const unsigned lineoff = ((unsigned)(c * M + wi * R) >> 4) << 7;
      const int g0 = (wi * CPI) & 7;
#pragma unroll
      for (int k = 0; k < CPI; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(Xb + tma_chunk_off(xbase, lineoff, g0 + k));
        a[2 * k] = make_float2(v.x, v.y);
        a[2 * k + 1] = make_float2(v.z, v.w);
      }
      Dft<R, true>::run(a);
    }
    __syncwarp();   // the lanes sharing this 128-byte line have read it
    if (act) {
      const int pb = sidx<LM>(c, wi * R);
#pragma unroll
      for (int m = 0; m < R; ++m) X[scomb<LM>(pb, m)] = a[m];
    }
  });
  __syncthreads();
  dim1_inv_stages_down<LM, NT, NS - 2>(X, ncols, tw, tid);
}

// Forward dim-1 FFT of `ncols` natural-order columns in X (must be synced) -> columns col0 .. col0+ncols-1 of the map.
// Thread 0 returns once the TMA unit has read the tile (the block may then exit).
template <int LM, int NT, bool WAIT = true>
ADMMTV_DI void dim1_fft_from_smem_tma(float2* X, int ncols, const CUtensorMap* map, int col0, const float2* __restrict__ tw, int tid) {
  constexpr int M = dim_len(LM), NS = plan_stages(M);
  dim1_fwd_stages_up<LM, NT, 0>(X, ncols, tw, tid);
  using St = Stage<M, NS - 1>;
  constexpr int R = St::R, CPI = R / 2;
  const unsigned xbase = smem_u32(X);
  char* Xb = reinterpret_cast<char*>(X);
  for_items_uniform<St::ITEMS, NT>(tid, ncols, [&](int wi, int c, bool act) {
    float2 a[R];
    if (act) {
      const int pb = sidx<LM>(c, wi * R);
#pragma unroll
      for (int m = 0; m < R; ++m) a[m] = X[scomb<LM>(pb, m)];
      Dft<R, false>::run(a);
    }
    __syncwarp();
    if (act) {
      const unsigned lineoff = ((unsigned)(c * M + wi * R) >> 4) << 7;
      const int g0 = (wi * CPI) & 7;
#pragma unroll
      for (int k = 0; k < CPI; ++k)
        *reinterpret_cast<float4*>(Xb + tma_chunk_off(xbase, lineoff, g0 + k)) = make_float4(a[2 * k].x, a[2 * k].y, a[2 * k + 1].x, a[2 * k + 1].y);
    }
  });
  fence_proxy_async();   // this thread's generic-proxy stores are visible to the async proxy
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < ncols; ++c) tma_store_3d(map, 0, 0, col0 + c, X + (size_t)c * M);
    tma_store_commit();
    if (WAIT) tma_store_wait_read();   // else the caller's thread 0 waits before it exits
  }
}
#endif  // !ADMMTV_EMU

// ------------------------------------------------------------------------------------------
// k_pack_fft1: user-layout planes -> pair-pack -> dim-1 FFT -> spectrum
//   MODE 0: src = y                                  (ops.jl:101 permute + first rfft pass)
//   MODE 1: src = xbar * act'(x_out); accumulates biasbar (deconv_admm.jl:222-224 pullback)
//   MODE 2: src = already pair-packed spatial data (b = H^T y -> first x-update input)
//   MODE 3: MODE 1 with the cotangent of the mean-squared-error loss formed on the fly: xbar = mse_scale (x_out - target);
//           also accumulates the loss (admmtv_backward_mse: no xbar array, no separate loss kernel)
// ------------------------------------------------------------------------------------------

template <int LM, int MODE, bool TMA>
ADMMTV_DI void pack_fft1_body(const PackArgs& A, const Dim1Tma* tm) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::CO;
  ADMMTV_DYN_SMEM(float2, X);
  // 1-D grid: block = (pair q, column tile), tiles fastest (no 65535 limit on the number of pairs)
  const int tid = threadIdx.x, N = A.N, ntile = (A.N + CO - 1) / CO;
  const int q = blockIdx.x / ntile;
  const int j0 = (blockIdx.x % ntile) * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  // MODE 1 reads the layer OUTPUT's cotangent, so it uses the output plane map
  constexpr bool COT = MODE == 1 || MODE == 3;   // reads the layer OUTPUT's cotangent
  const long ia = COT ? pm_out(A.pm, q, 0) : pm_in(A.pm, q, 0);
  const long ib = COT ? pm_out(A.pm, q, 1) : pm_in(A.pm, q, 1);
  const bool has_b = ib >= 0;
  const float* srcp = MODE == 3 ? A.target : A.src;
  const float* pa = srcp + (size_t)ia * plane;
  const float* pb = srcp + (size_t)(has_b ? ib : 0) * plane;
  double bsum = 0.0, lsum = 0.0;
  // 4 rows per thread: float4 loads from each plane (M is a multiple of 32)
  for (int e = tid; e < nout * (M / 4); e += NT) {
    const int c = e / (M / 4), i = (e % (M / 4)) * 4;
    const size_t off = (size_t)(j0 + c) * M + i;
    if (MODE == 2) {
      const float4 lo = *reinterpret_cast<const float4*>(A.src_packed + (size_t)q * plane + off);
      const float4 hi = *reinterpret_cast<const float4*>(A.src_packed + (size_t)q * plane + off + 2);
      X[sidx<LM>(c, i)] = make_float2(lo.x, lo.y);
      X[sidx<LM>(c, i + 1)] = make_float2(lo.z, lo.w);
      X[sidx<LM>(c, i + 2)] = make_float2(hi.x, hi.y);
      X[sidx<LM>(c, i + 3)] = make_float2(hi.z, hi.w);
      continue;
    }
    const float4 a4 = *reinterpret_cast<const float4*>(pa + off);
    const float4 b4 = has_b ? *reinterpret_cast<const float4*>(pb + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    float va[4] = {a4.x, a4.y, a4.z, a4.w}, vb[4] = {b4.x, b4.y, b4.z, b4.w};
    if (MODE == 0 && A.packed_out) {
      float4* po = reinterpret_cast<float4*>(A.packed_out + (size_t)q * plane + off);
      po[0] = make_float4(va[0], vb[0], va[1], vb[1]);
      po[1] = make_float4(va[2], vb[2], va[3], vb[3]);
    }
    if (COT) {
      const float4 oa = *reinterpret_cast<const float4*>(A.xout + (size_t)ia * plane + off);
      const float4 ob = has_b ? *reinterpret_cast<const float4*>(A.xout + (size_t)ib * plane + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float xa[4] = {oa.x, oa.y, oa.z, oa.w}, xb[4] = {ob.x, ob.y, ob.z, ob.w};
      if (MODE == 3) {
        float l = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float da = xa[t] - va[t], db = has_b ? xb[t] - vb[t] : 0.f;
          l += da * da + db * db;
          va[t] = A.mse_scale * da;
          vb[t] = A.mse_scale * db;
        }
        lsum += (double)l;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        va[t] *= act_grad_from_out(xa[t], A.act);
        if (has_b) vb[t] *= act_grad_from_out(xb[t], A.act);
        bsum += (double)va[t] + (double)vb[t];
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) X[sidx<LM>(c, i + t)] = make_float2(va[t], vb[t]);
  }
  if (COT && A.bias_acc) {
    const double tot = block_sum(bsum);
    if (tid == 0) atomicAdd(A.bias_acc + (size_t)A.acc_stride * (q / A.pm.Qg), tot);  // biasbar slot of this pair's group
  }
  if (MODE == 3) {
    const double tot = block_sum(lsum);
    if (tid == 0) atomicAdd(A.loss_acc, tot);
  }
  __syncthreads();
  float2* sq = A.spec + (size_t)q * plane;
#ifndef ADMMTV_EMU
  if constexpr (TMA) dim1_fft_from_smem_tma<LM, NT>(X, nout, &tm->out, q * N + j0, A.twM, tid);
  else
#endif
    dim1_fft_from_smem<LM, NT>(X, nout, [&](int c) { return sq + (size_t)(j0 + c) * M; }, A.twM, tid);
}
template <int LM, int MODE>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_pack_fft1(PackArgs A) {
  pack_fft1_body<LM, MODE, false>(A, nullptr);
}
#ifndef ADMMTV_EMU
template <int LM, int MODE>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_pack_fft1_tma(PackArgs A, const __grid_constant__ Dim1Tma tm) {
  pack_fft1_body<LM, MODE, true>(A, &tm);
}
#endif

// ------------------------------------------------------------------------------------------
// k_dim1_out: spectrum -> dim-1 IFFT -> spatial
//   MODE 0: pair-packed float2 out (b = H^T y)
//   MODE 1: user-layout planes, + bias, activation           (deconv_admm.jl:222-224, ops.jl:175)
// ------------------------------------------------------------------------------------------

template <int LM, int MODE, bool TMA>
ADMMTV_DI void dim1_out_body(const OutArgs& A, const Dim1Tma* tm) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::CO;
  ADMMTV_DYN_SMEM(float2, X);
  // 1-D grid: block = (pair q, column tile), tiles fastest (no 65535 limit on the number of pairs)
  const int tid = threadIdx.x, N = A.N, ntile = (A.N + CO - 1) / CO;
  const int q = blockIdx.x / ntile;
  const int j0 = (blockIdx.x % ntile) * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  const float2* sq = A.spec + (size_t)q * plane;
#ifndef ADMMTV_EMU
  __shared__ unsigned long long tbar[1];
  if constexpr (TMA) dim1_ifft_to_smem_tma<LM, NT>(X, nout, &tm->in, [&](int c) { return q * N + j0 + c; }, tbar, A.twM, tid);
  else
#endif
    dim1_ifft_to_smem<LM, NT>(X, nout, [&](int c) { return sq + (size_t)(j0 + c) * M; }, A.twM, tid);
  if (MODE == 0) {
    float2* dst = A.packed + (size_t)q * plane;
    const float sc = A.scale;
    for (int e = tid; e < nout * (M / 2); e += NT) {
      const int c = e / (M / 2), i = (e % (M / 2)) * 2;
      const float2 v0 = X[sidx<LM>(c, i)], v1 = X[sidx<LM>(c, i + 1)];
      *reinterpret_cast<float4*>(dst + (size_t)(j0 + c) * M + i) = make_float4(sc * v0.x, sc * v0.y, sc * v1.x, sc * v1.y);
    }
  } else if (MODE == 2) {
    // cotangent of the INPUT: input plane map; groups sharing y accumulate (ybar zeroed by the host)
    const long ia = pm_in(A.pm, q, 0), ib = pm_in(A.pm, q, 1);
    const bool shared = A.pm.G > 1 && A.pm.in_gstride == 0;
    const float sc = A.scale;
    for (int e = tid; e < nout * M; e += NT) {
      const int c = e / M, i = e % M;
      const float2 v = cscale(X[sidx<LM>(c, i)], sc);
      const size_t off = (size_t)(j0 + c) * M + i;
      if (shared) {
        atomicAdd(A.planes + (size_t)ia * plane + off, v.x);
        if (ib >= 0) atomicAdd(A.planes + (size_t)ib * plane + off, v.y);
      } else {
        A.planes[(size_t)ia * plane + off] = v.x;
        if (ib >= 0) A.planes[(size_t)ib * plane + off] = v.y;
      }
    }
  } else {
    const long ia = pm_out(A.pm, q, 0), ib = pm_out(A.pm, q, 1);
    const bool has_b = ib >= 0;
    const float bias = A.bias ? A.bias[q / A.pm.Qg] : 0.f;
    float* pa = A.planes + (size_t)ia * plane;
    float* pb = A.planes + (size_t)(has_b ? ib : 0) * plane;
    for (int e = tid; e < nout * (M / 4); e += NT) {
      const int c = e / (M / 4), i = (e % (M / 4)) * 4;
      float2 v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) v[t] = X[sidx<LM>(c, i + t)];
      const size_t off = (size_t)(j0 + c) * M + i;
      *reinterpret_cast<float4*>(pa + off) = make_float4(act_apply(v[0].x + bias, A.act), act_apply(v[1].x + bias, A.act),
                                                          act_apply(v[2].x + bias, A.act), act_apply(v[3].x + bias, A.act));
      if (has_b)
        *reinterpret_cast<float4*>(pb + off) = make_float4(act_apply(v[0].y + bias, A.act), act_apply(v[1].y + bias, A.act),
                                                            act_apply(v[2].y + bias, A.act), act_apply(v[3].y + bias, A.act));
    }
  }
}
template <int LM, int MODE>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_dim1_out(OutArgs A) {
  dim1_out_body<LM, MODE, false>(A, nullptr);
}
#ifndef ADMMTV_EMU
template <int LM, int MODE>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_dim1_out_tma(OutArgs A, const __grid_constant__ Dim1Tma tm) {
  dim1_out_body<LM, MODE, true>(A, &tm);
}
#endif

// ------------------------------------------------------------------------------------------
// k_dim1_fwd: the fused per-iteration kernel (anisotropic)
//   dim-1 IFFT -> x_k                                             (ops.jl:168 irfft, last pass)
//   v_k = D x_k + u_{k-1},  u_{k-1} = v_{k-1} - ST(v_{k-1}, tau)    (ops.jl:169,171)
//   z_k = ST(v_k, tau) ; u_k = v_k - z_k                           (ops.jl:171,173)
//   r_{k+1} = H^T y + rho D^T (z_k - u_k)                          (ops.jl:168, next iteration)
//   dim-1 FFT of r_{k+1}                                           (ops.jl:168 rfft, first pass)
// The only state written is v_k (z_k, u_k are recomputed from it), which is also the backward
// checkpoint.
// ------------------------------------------------------------------------------------------

struct Shrunk {
  float2 u, w;  // u = v - z ; w = z - u
};
ADMMTV_DI float st1(float v, float tau) { return v - fminf(fmaxf(v, -tau), tau); }  // ST, ops.jl:9
ADMMTV_DI Shrunk shrink_aniso(float2 v, float tau) {
  const float2 z = make_float2(st1(v.x, tau), st1(v.y, tau));
  Shrunk s;
  s.u = csub(v, z);
  s.w = csub(z, s.u);
  return s;
}

template <int RPT>
ADMMTV_DI void load_rows(const float2* __restrict__ p, float2* out) {
  if constexpr (RPT % 2 == 0) load_contig<RPT>(p, out);
  else {
#pragma unroll
    for (int r = 0; r < RPT; ++r) out[r] = p[r];
  }
}
template <int RPT>
ADMMTV_DI void store_rows(float2* __restrict__ p, const float2* v) {
  if constexpr (RPT % 2 == 0) store_contig<RPT>(p, v);
  else {
#pragma unroll
    for (int r = 0; r < RPT; ++r) p[r] = v[r];
  }
}

template <int RPT>
ADMMTV_DI void load_rows_f(const float* __restrict__ p, float* out) {
  if constexpr (RPT % 4 == 0) {
#pragma unroll
    for (int r = 0; r < RPT / 4; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(p + 4 * r);
      out[4 * r] = v.x; out[4 * r + 1] = v.y; out[4 * r + 2] = v.z; out[4 * r + 3] = v.w;
    }
  } else if constexpr (RPT % 2 == 0) {
#pragma unroll
    for (int r = 0; r < RPT / 2; ++r) {
      const float2 v = *reinterpret_cast<const float2*>(p + 2 * r);
      out[2 * r] = v.x; out[2 * r + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int r = 0; r < RPT; ++r) out[r] = p[r];
  }
}

template <int RPT>
ADMMTV_DI void store_rows_f(float* __restrict__ p, const float* v) {
  if constexpr (RPT % 4 == 0) {
#pragma unroll
    for (int r = 0; r < RPT / 4; ++r) *reinterpret_cast<float4*>(p + 4 * r) = make_float4(v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);
  } else if constexpr (RPT % 2 == 0) {
#pragma unroll
    for (int r = 0; r < RPT / 2; ++r) *reinterpret_cast<float2*>(p + 2 * r) = make_float2(v[2 * r], v[2 * r + 1]);
  } else {
#pragma unroll
    for (int r = 0; r < RPT; ++r) p[r] = v[r];
  }
}

// isotropic block thresholding (ops.jl:6,10): one scale per pixel, s = max(1 - tau/n, 0)
ADMMTV_DI float iso_scale(float nsq, float tau) {
  const float n = sqrtf(nsq);
  return n > 0.f ? fmaxf(1.f - tau / n, 0.f) : 0.f;  // n = 0: max(-Inf,0)*0 = 0
}
ADMMTV_DI Shrunk shrink_iso(float2 v, float s) {
  const float2 z = make_float2(s * v.x, s * v.y);
  Shrunk r;
  r.u = csub(v, z);
  r.w = csub(z, r.u);
  return r;
}

// MODE 0: the fused anisotropic kernel described above.
// MODE 1: isotropic pass B -- v_k (A.vprev) and the per-pixel shrink scale s_k = max(1 - tau/n_k, 0) (A.nsq,
//         computed by k_iso_scale) are given; no IFFT, no state write: w = (2s-1) v, r = b + rho D^T w, FFT.
// MODE 2: isotropic pass A -- dim-1 IFFT -> x_k ; v_k = D x_k + u_{k-1} with u_{k-1} = (1 - s_{k-1}) v_{k-1}
//         (A.vprev, A.nsq = s_{k-1}) ; store v_k ; A.nsq_out[q][pixel] = this pair's share of |v_k|^2 (plain stores:
//         k_iso_scale adds the pairs of a group in a fixed order, so the norm is bit-reproducible) ; no FFT.
template <int LM, bool HAS_VPREV, int MODE, bool TMA>
ADMMTV_DI void dim1_fwd_body(const Dim1FwdArgs& A, const Dim1Tma* tm) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Dim1FwdCfg<LM, MODE>::NT, RB = Dim1FwdCfg<LM, MODE>::RB, CO = Cfg::CO, RPT = Cfg::RPT, CHUNK = Cfg::CHUNK;
  static_assert(RPT * NT * RB == M, "row blocks tile the column");
  // x at the rows either side of the row-block boundary (rows 0 and M/2 - 1) of every tile column: the second sweep needs
  // them after the first one has overwritten its half of the tile with r_{k+1}
  __shared__ float2 xedge[RB > 1 ? 2 : 1][RB > 1 ? Cfg::TC : 1];
  static_assert(CO % CHUNK == 0, "chunking must divide the tile");
  ADMMTV_DYN_SMEM(float2, X);
  // 1-D grid: block = (pair q, column tile), tiles fastest (no 65535 limit on the number of pairs)
  const int tid = threadIdx.x, N = A.N, ntile = (A.N + CO - 1) / CO;
  const int bid = dim1_bid();
  const int q = bid / ntile;
  const int j0 = (bid % ntile) * CO;
  const int nout = min(CO, N - j0);  // N and CO are powers of two: nout % CHUNK == 0
  const size_t plane = (size_t)N * M;
  const float2* sin_q = A.spec_in + (size_t)q * plane;
  auto jcol = [&](int c) {
    int j = j0 - 1 + c;
    if (j < 0) j += N;
    if (j >= N) j -= N;
    return j;
  };

  // stencil inputs of this tile -> L2 while the IFFT passes run (one bulk prefetch per array)
  {
    const unsigned colb = (unsigned)(M * sizeof(float2));
    if (HAS_VPREV && tid == 0) {
      l2_prefetch_bulk(A.vprev + ((size_t)q * 2 + 0) * plane + (size_t)j0 * M, (unsigned)nout * colb);
      l2_prefetch_bulk(A.vprev + ((size_t)q * 2 + 0) * plane + (size_t)jcol(nout + 1) * M, colb);
    }
    if (HAS_VPREV && tid == 32 % NT) l2_prefetch_bulk(A.vprev + ((size_t)q * 2 + 1) * plane + (size_t)j0 * M, (unsigned)nout * colb);
    if (MODE != 2 && tid == 64 % NT) l2_prefetch_bulk(A.bpk + (size_t)q * plane + (size_t)j0 * M, (unsigned)nout * colb);
    // the spectrum tile of the block that will be scheduled onto this SM slot when a resident block retires
    if (MODE == 0 && ADMMTV_PF_NEXT > 0 && tid == 96 % NT) {
      const long nid = (long)bid + ADMMTV_PF_NEXT;
      if (nid < (long)gridDim.x) {
        const int nq = (int)(nid / ntile), nj0 = (int)(nid % ntile) * CO;
        const int ncol = min(CO, N - nj0);
        l2_prefetch_bulk(A.spec_in + (size_t)nq * plane + (size_t)nj0 * M, (unsigned)ncol * colb);
      }
    }
  }
  // 1. x_k for columns j0-1 .. j0+nout (one halo column each side)
#ifndef ADMMTV_EMU
  __shared__ unsigned long long tbar[1];
  if constexpr (TMA) {
    if (MODE != 1) dim1_ifft_to_smem_tma<LM, NT>(X, MODE == 2 ? nout + 1 : nout + 2, &tm->in, [&](int c) { return q * N + jcol(c); }, tbar, A.twM, tid);
  } else
#endif
  {
    if (MODE == 0) dim1_ifft_to_smem<LM, NT>(X, nout + 2, [&](int c) { return sin_q + (size_t)jcol(c) * M; }, A.twM, tid);
    if (MODE == 2) dim1_ifft_to_smem<LM, NT>(X, nout + 1, [&](int c) { return sin_q + (size_t)jcol(c) * M; }, A.twM, tid);
  }

  // 2. stencil sweep: this thread owns rows i0 .. i0+RPT-1 of every column
  const int grp = q / A.Qg;
  const float rho = A.rho[grp * A.PS + A.in];                                     // rho of r_{k+1}
  const float tau = A.lambda[grp * A.PS + A.ic] / A.rho[grp * A.PS + A.ic];       // tau of v_k      (ops.jl:102)
  const float tau_p = A.lambda[grp * A.PS + A.ip] / A.rho[grp * A.PS + A.ip];     // tau of v_{k-1}
  if constexpr (RB > 1) {
    if (tid < nout + 2) {
      xedge[0][tid] = X[sidx<LM>(tid, 0)];
      xedge[1][tid] = X[sidx<LM>(tid, M / 2 - 1)];
    }   // read in the second sweep only: the first sweep's chunk barriers order these stores before it
  }
  const float* nsq_g = MODE != 0 ? A.nsq + (size_t)grp * plane : nullptr;
  float* nsq_o = MODE == 2 ? A.nsq_out + (size_t)q * plane : nullptr;   // per-pair partial sums
  const float2* vp1 = A.vprev + ((size_t)q * 2 + 0) * plane;
  const float2* vp2 = A.vprev + ((size_t)q * 2 + 1) * plane;
  float2* vn1 = A.vnew + ((size_t)q * 2 + 0) * plane;
  float2* vn2 = A.vnew + ((size_t)q * 2 + 1) * plane;
  const float2* bq = A.bpk + (size_t)q * plane;

#pragma unroll 1
  for (int rb = 0; rb < RB; ++rb) {
  const int i0 = tid * RPT + rb * (M / RB);
  float2 w1c[RPT];   // MODE 0/1: w1 of the current column ; MODE 2: v1 of the current column
  if (MODE == 2) {
    const int j = jcol(1);
    float2 up[RPT];
    float nn[RPT];
    if (HAS_VPREV) {
      load_rows<RPT>(vp1 + (size_t)j * M + i0, up);
      load_rows_f<RPT>(nsq_g + (size_t)j * M + i0, nn);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      float2 v = csub(X[sidx<LM>(1, i0 + r)], X[sidx<LM>(0, i0 + r)]);
      if (HAS_VPREV) v = cadd(v, shrink_iso(up[r], nn[r]).u);
      w1c[r] = v;
    }
    store_rows<RPT>(vn1 + (size_t)j * M + i0, w1c);
  } else if (MODE == 1) {
    const int j = jcol(1);
    float2 vv[RPT];
    float nn[RPT];
    load_rows<RPT>(vp1 + (size_t)j * M + i0, vv);
    load_rows_f<RPT>(nsq_g + (size_t)j * M + i0, nn);
#pragma unroll
    for (int r = 0; r < RPT; ++r) w1c[r] = shrink_iso(vv[r], nn[r]).w;
  } else {
    const int j = jcol(1);
    float2 up[RPT], vst[RPT];
    if (HAS_VPREV) load_rows<RPT>(vp1 + (size_t)j * M + i0, up);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float2 xa = X[sidx<LM>(0, i0 + r)], xb = X[sidx<LM>(1, i0 + r)];
      float2 v = csub(xb, xa);
      if (HAS_VPREV) v = cadd(v, shrink_aniso(up[r], tau_p).u);
      vst[r] = v;
      w1c[r] = shrink_aniso(v, tau).w;
    }
    store_rows<RPT>(vn1 + (size_t)j * M + i0, vst);
  }

  for (int c = 1; c <= nout; c += CHUNK) {
    // (a) every global load of the chunk is issued before any of them is consumed
    float2 g1[CHUNK][RPT], g2[CHUNK][RPT + 1], gb[CHUNK][RPT];
    float n1[MODE != 0 ? CHUNK : 1][RPT], n2[MODE != 0 ? CHUNK : 1][RPT + 1];
#pragma unroll
    for (int cc = 0; cc < CHUNK; ++cc) {
      const int j = jcol(c + cc), jn = jcol(c + cc + 1);
      if (MODE == 1 || (MODE == 2 && HAS_VPREV)) {
        load_rows_f<RPT>(nsq_g + (size_t)jn * M + i0, n1[cc]);
        load_rows_f<RPT>(nsq_g + (size_t)j * M + i0, n2[cc]);
        if (MODE == 1) n2[cc][RPT] = nsq_g[(size_t)j * M + wrapm<M>(i0 + RPT)];
      }
      if (MODE == 2 && HAS_VPREV) {
        load_rows<RPT>(vp1 + (size_t)jn * M + i0, g1[cc]);
        load_rows<RPT>(vp2 + (size_t)j * M + i0, g2[cc]);
      } else if (HAS_VPREV) {
        load_rows<RPT>(vp1 + (size_t)jn * M + i0, g1[cc]);
        load_rows<RPT>(vp2 + (size_t)j * M + i0, g2[cc]);
        g2[cc][RPT] = vp2[(size_t)j * M + wrapm<M>(i0 + RPT)];
      }
      if (MODE != 2) load_rows<RPT>(bq + (size_t)j * M + i0, gb[cc]);
    }
    // (b) stencil
    float2 rr[CHUNK][RPT];
#pragma unroll
    for (int cc = 0; cc < CHUNK; ++cc) {
      const int col = c + cc;
      const int j = jcol(col), jn = jcol(col + 1);
      float2 w1n[RPT];
      float2 w2[RPT + 1];
      if (MODE == 2) {
        float2 xc[RPT + 1];  // rows i0-1 .. i0+RPT-1 of column col
        xc[0] = X[sidx<LM>(col, wrapm<M>(i0 - 1))];
#pragma unroll
        for (int r = 0; r < RPT; ++r) xc[r + 1] = X[sidx<LM>(col, i0 + r)];
        if (col < nout) {   // channel 1 of the next own column
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            float2 v = csub(X[sidx<LM>(col + 1, i0 + r)], xc[r + 1]);
            if (HAS_VPREV) v = cadd(v, shrink_iso(g1[cc][r], n1[cc][r]).u);
            w1n[r] = v;
          }
          store_rows<RPT>(vn1 + (size_t)jn * M + i0, w1n);
        }
        float2 vst[RPT];
        float sq[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          float2 v = csub(xc[r + 1], xc[r]);
          if (HAS_VPREV) v = cadd(v, shrink_iso(g2[cc][r], n2[cc][r]).u);
          vst[r] = v;
          sq[r] = w1c[r].x * w1c[r].x + w1c[r].y * w1c[r].y + v.x * v.x + v.y * v.y;
          w1c[r] = w1n[r];
        }
        store_rows<RPT>(vn2 + (size_t)j * M + i0, vst);
        store_rows_f<RPT>(nsq_o + (size_t)j * M + i0, sq);
        continue;
      }
      if (MODE == 1) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) w1n[r] = shrink_iso(g1[cc][r], n1[cc][r]).w;
#pragma unroll
        for (int r = 0; r <= RPT; ++r) w2[r] = shrink_iso(g2[cc][r], n2[cc][r]).w;
      } else {
      float2 xc[RPT + 2];  // rows i0-1 .. i0+RPT of column col
      xc[0] = X[sidx<LM>(col, wrapm<M>(i0 - 1))];
#pragma unroll
      for (int r = 0; r < RPT; ++r) xc[r + 1] = X[sidx<LM>(col, i0 + r)];
      xc[RPT + 1] = X[sidx<LM>(col, wrapm<M>(i0 + RPT))];
      if constexpr (RB > 1) {   // second sweep: the first one has replaced rows < M/2 of the slots behind it
        if (rb == 1 && tid == 0) xc[0] = xedge[1][col];
        if (rb == 1 && tid == NT - 1) xc[RPT + 1] = xedge[0][col];
      }

      // channel 1 (dim-2 difference) at column col+1
      {
        float2 vst[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          float2 v = csub(X[sidx<LM>(col + 1, i0 + r)], xc[r + 1]);
          if (HAS_VPREV) v = cadd(v, shrink_aniso(g1[cc][r], tau_p).u);
          vst[r] = v;
          w1n[r] = shrink_aniso(v, tau).w;
        }
        if (col + 1 <= nout) store_rows<RPT>(vn1 + (size_t)jn * M + i0, vst);
      }
      // channel 2 (dim-1 difference) at column col, rows i0 .. i0+RPT (the last is the
      // neighbour's first row, recomputed here instead of exchanged)
      {
        float2 vst[RPT];
#pragma unroll
        for (int r = 0; r <= RPT; ++r) {
          float2 v = csub(xc[r + 1], xc[r]);
          if (HAS_VPREV) v = cadd(v, shrink_aniso(g2[cc][r], tau_p).u);
          if (r < RPT) vst[r] = v;
          w2[r] = shrink_aniso(v, tau).w;
        }
        store_rows<RPT>(vn2 + (size_t)j * M + i0, vst);
      }
      }  // MODE 0
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float2 dt = cadd(csub(w1c[r], w1n[r]), csub(w2[r], w2[r + 1]));  // D^T(z-u)
        rr[cc][r] = make_float2(gb[cc][r].x + rho * dt.x, gb[cc][r].y + rho * dt.y);
        w1c[r] = w1n[r];
      }
    }
    if (MODE == 2) continue;   // X is only read in pass A
    __syncthreads();  // every thread is done reading columns <= c+CHUNK-1 of this chunk
#pragma unroll
    for (int cc = 0; cc < CHUNK; ++cc)
#pragma unroll
      for (int r = 0; r < RPT; ++r) X[sidx<LM>(c + cc - 1, i0 + r)] = rr[cc][r];  // r column -> slot col-1
  }
  }  // row blocks
  if (MODE == 2) return;
  __syncthreads();

  // 3. dim-1 FFT of r_{k+1}: slot s holds output column j0+s
  float2* sout_q = A.spec_out + (size_t)q * plane;
#ifndef ADMMTV_EMU
  if constexpr (TMA) dim1_fft_from_smem_tma<LM, NT>(X, nout, &tm->out, q * N + j0, A.twM, tid);
  else
#endif
    dim1_fft_from_smem<LM, NT>(X, nout, [&](int c) { return sout_q + (size_t)(j0 + c) * M; }, A.twM, tid);
}
template <int LM, bool HAS_VPREV, int MODE = 0>
__global__ void __launch_bounds__(Dim1FwdCfg<LM, MODE>::NT, Dim1FwdCfg<LM, MODE>::MINB) k_dim1_fwd(Dim1FwdArgs A) {
  dim1_fwd_body<LM, HAS_VPREV, MODE, false>(A, nullptr);
}
#ifndef ADMMTV_EMU
// the iteration kernel with TMA column loads / stores of the spectra (Dim1Tma: maps of spec_in / spec_out; isotropic pass A
// only loads, pass B only stores)
template <int LM, bool HAS_VPREV, int MODE = 0>
__global__ void __launch_bounds__(Dim1FwdCfg<LM, MODE>::NT, Dim1FwdCfg<LM, MODE>::MINB) k_dim1_fwd_tma(Dim1FwdArgs A, const __grid_constant__ Dim1Tma tm) {
  dim1_fwd_body<LM, HAS_VPREV, MODE, true>(A, &tm);
}
#endif

// ------------------------------------------------------------------------------------------
// k_dim2: dim-2 (strided) pass over a tile of TR contiguous rows x all N columns
//   forward FFT along dim 2 -> [save Z] -> [accumulate conj(Z) Z2] -> multiply by table ->
//   inverse FFT along dim 2                                  (ops.jl:168  C .* rfft(...))
// ------------------------------------------------------------------------------------------

constexpr int kP2 = 1;   // plan id of the dim-2 transforms (fft_core.cuh)

template <int LN>
struct Dim2Cfg {
  static constexpr int N = dim_len(LN);
  static constexpr int TR = !is_pow2(N) ? (N <= 640 ? 16 : 8)
                                        : (LN <= 6 ? 16 : (LN == 7 ? ADMMTV_TR7 : (LN == 8 ? ADMMTV_TR8 : (LN == 9 ? ADMMTV_TR9 : (LN == 10 ? 8 : (LN == 11 ? ADMMTV_TR11 : 4))))));
  // block size = work items of the widest pass (row pairs x N / largest radix), within [32, 512]
  static constexpr int ITEMS_MAX = (TR / 2) * (N / plan_radix(N, 0, kP2));
  static constexpr int NT_AUTO = ITEMS_MAX < 32 ? 32 : (ITEMS_MAX > ADMMTV_NT2_MAX ? ADMMTV_NT2_MAX : ITEMS_MAX);
  static constexpr int NT = LN == 9 ? ADMMTV_NT2 : NT_AUTO;
  static constexpr int MINB = LN == 9 ? ADMMTV_MINB2 : (LN == 11 ? ADMMTV_MINB2_11 : 1);
  static constexpr size_t SMEM = (size_t)N * TR * sizeof(float2);
  // Tiles narrower than 16 rows (a column = 64 or 32 bytes of the 128-byte bank span): the last plan stage reads columns
  // R apart, which land on the same banks (ncu, N = 2048: 20 % of the shared-memory wavefronts are conflicts at TR = 8,
  // 43 % at TR = 4).  Column c is therefore stored at c ^ ((c >> log2 R_last) & (16/TR - 1)): the R-apart columns of a
  // warp's work items rotate through the 16/TR column slots of a bank span, and every other stage (whose lanes hold
  // consecutive columns with constant high bits) sees a permutation inside aligned groups, i.e. stays conflict-free.
  static constexpr bool SWZ = ADMMTV_D2_SWZ && is_pow2(N) && TR < 16;
};
template <int LN, int TR, bool SWZ>
ADMMTV_DI int d2col(int c) {
  if constexpr (!SWZ) return c;
  else {
    constexpr int N = dim_len(LN), RL = plan_radix(N, plan_stages(N, kP2) - 1, kP2);
    constexpr int SH = RL == 32 ? 5 : (RL == 16 ? 4 : (RL == 8 ? 3 : 2));
    return c ^ ((c >> SH) & (16 / TR - 1));
  }
}

template <int LN, int S, bool INV, int NTX = Dim2Cfg<LN>::NT, int TRX = Dim2Cfg<LN>::TR, bool SWZ = false>
ADMMTV_DI void dim2_smem_stage(float2* tile, const float2* __restrict__ tw, int tid) {
  using Cfg = Dim2Cfg<LN>;
  constexpr int N = Cfg::N, TR = TRX, NT = NTX, RP = TR / 2;
  using St = Stage<N, S, kP2>;
  for (int item = tid; item < RP * St::ITEMS; item += NT) {
    const int rp = item % RP, wi = item / RP;
    float2 p[St::R];
    stage_twiddles<N, S, INV, kP2>(wi, tw, p);
    float2 a0[St::R], a1[St::R];
    const int base = St::base(wi);
#pragma unroll
    for (int m = 0; m < St::R; ++m) {
      const float4 v = *reinterpret_cast<const float4*>(tile + d2col<LN, TR, SWZ>(base + m * St::STRIDE) * TR + 2 * rp);
      a0[m] = make_float2(v.x, v.y);
      a1[m] = make_float2(v.z, v.w);
    }
    if (INV) {
      stage_inv<N, S, kP2>(a0, p);
      stage_inv<N, S, kP2>(a1, p);
    } else {
      stage_fwd<N, S, kP2>(a0, p);
      stage_fwd<N, S, kP2>(a1, p);
    }
#pragma unroll
    for (int m = 0; m < St::R; ++m)
      *reinterpret_cast<float4*>(tile + d2col<LN, TR, SWZ>(base + m * St::STRIDE) * TR + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
  }
}
template <int LN, int S, int NTX = Dim2Cfg<LN>::NT>
ADMMTV_DI void dim2_fwd_mid(float2* tile, const float2* __restrict__ tw, int tid) {
  if constexpr (S < plan_stages(dim_len(LN), kP2) - 1) {
    dim2_smem_stage<LN, S, false, NTX, Dim2Cfg<LN>::TR, Dim2Cfg<LN>::SWZ>(tile, tw, tid);
    __syncthreads();
    dim2_fwd_mid<LN, S + 1, NTX>(tile, tw, tid);
  }
}
template <int LN, int S, int NTX = Dim2Cfg<LN>::NT>
ADMMTV_DI void dim2_inv_mid(float2* tile, const float2* __restrict__ tw, int tid) {
  if constexpr (S >= 1) {
    dim2_smem_stage<LN, S, true, NTX, Dim2Cfg<LN>::TR, Dim2Cfg<LN>::SWZ>(tile, tw, tid);
    __syncthreads();
    dim2_inv_mid<LN, S - 1, NTX>(tile, tw, tid);
  }
}

// Per-thread L2 prefetch of one 128-byte line (the next tile's first-pass loads hit L2).
ADMMTV_DI void l2_prefetch_line(const void* p) {
#if !defined(ADMMTV_EMU) && ADMMTV_PREFETCH
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// The last forward stage fused with [save Z] -> [accumulate conj(Z) Z2] -> spectral multiply -> first inverse stage,
// over the whole tile.  Two mappings: radix <= 16 works on ROW PAIRS (float4 shared-memory accesses, both rows share the
// index arithmetic); radix 32 works on single rows (32 complex values are all the registers a thread has).  Each tile
// element is owned by exactly one thread in either mapping (dim2_fused_flush relies on it).
template <int LN, int MUL, bool SAVE_Z, int ACC, bool FWD_ONLY, int NT, int TR, bool SWZ = false>
ADMMTV_DI void dim2_fused_stage(float2* tile, float* gsm, const Dim2Args& A, size_t qoff, size_t toff, int i0, int tid) {
  constexpr int N = dim_len(LN), NS = plan_stages(N, kP2), RP = TR / 2;
  using StL = Stage<N, NS - 1, kP2>;
  static_assert(StL::STRIDE == 1, "last plan stage must be contiguous");
  constexpr bool SMACC = ACC == 1;
  const int M = A.M;
  if constexpr (StL::R <= 16) {
#pragma unroll 1
    for (int item = tid; item < RP * StL::ITEMS; item += NT) {
      const int rp = item % RP, wi = item / RP;
      float2 a0[StL::R], a1[StL::R];
      float2 cpre[(ADMMTV_D2_CPRE && MUL == 0 && !FWD_ONLY) ? StL::R : 1];
      if (ADMMTV_D2_CPRE && MUL == 0 && !FWD_ONLY) {   // table values in flight while the butterflies run
#pragma unroll
        for (int m = 0; m < StL::R; ++m)
          cpre[m] = *reinterpret_cast<const float2*>(A.ctab + toff + (size_t)(wi * StL::R + m) * M + i0 + 2 * rp);
      }
#pragma unroll
      for (int m = 0; m < StL::R; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(tile + d2col<LN, TR, SWZ>(wi * StL::R + m) * TR + 2 * rp);
        a0[m] = make_float2(v.x, v.y);
        a1[m] = make_float2(v.z, v.w);
      }
      Dft<StL::R, false>::run(a0);
      Dft<StL::R, false>::run(a1);
#pragma unroll
      for (int m = 0; m < StL::R; ++m) {
        const size_t g = (size_t)(wi * StL::R + m) * M + i0 + 2 * rp;  // table / spectrum offset
        if (SAVE_Z || FWD_ONLY) {
          float2* zs = FWD_ONLY ? A.out : A.zsave;
          *reinterpret_cast<float4*>(zs + qoff + g) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
        }
        if (ACC != 0) {
          const float4 z2 = *reinterpret_cast<const float4*>(A.z2 + qoff + g);
          // conj(Z) * Z2
          const float re0 = a0[m].x * z2.x + a0[m].y * z2.y, im0 = a0[m].x * z2.y - a0[m].y * z2.x;
          const float re1 = a1[m].x * z2.z + a1[m].y * z2.w, im1 = a1[m].x * z2.w - a1[m].y * z2.z;
          if (SMACC) {
            float2* gp = reinterpret_cast<float2*>(gsm + (wi * StL::R + m) * TR + 2 * rp);
            float2 gv = *gp;
            gv.x += re0;
            gv.y += re1;
            *gp = gv;
          } else {
            float* pa = reinterpret_cast<float*>(A.pacc + toff + g);
            atomicAdd(pa, re0);
            atomicAdd(pa + 1, im0);
            atomicAdd(pa + 2, re1);
            atomicAdd(pa + 3, im1);
          }
        }
        if (!FWD_ONLY) {
          if (MUL == 0) {
            const float2 cc = ADMMTV_D2_CPRE ? cpre[m] : *reinterpret_cast<const float2*>(A.ctab + toff + g);
            a0[m] = cscale(a0[m], cc.x);
            a1[m] = cscale(a1[m], cc.y);
          } else {
            const float4 kk = *reinterpret_cast<const float4*>(A.ktab + toff + g);
            const float sgn = MUL == 2 ? -1.f : 1.f;
            a0[m] = cmul(a0[m], make_float2(kk.x, sgn * kk.y));
            a1[m] = cmul(a1[m], make_float2(kk.z, sgn * kk.w));
          }
        }
      }
      if (!FWD_ONLY) {
        Dft<StL::R, true>::run(a0);
        Dft<StL::R, true>::run(a1);
#pragma unroll
        for (int m = 0; m < StL::R; ++m)
          *reinterpret_cast<float4*>(tile + d2col<LN, TR, SWZ>(wi * StL::R + m) * TR + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
      }
    }
  } else {
    static_assert(!SWZ, "the single-row mapping (radix 32) is only used with 16-row tiles");
    // single rows: lanes run along the TR contiguous rows of a column (conflict-free 8-byte accesses)
#pragma unroll 1
    for (int item = tid; item < TR * StL::ITEMS; item += NT) {
      const int r = item % TR, wi = item / TR;
      float2 a[StL::R];
#pragma unroll
      for (int m = 0; m < StL::R; ++m) a[m] = tile[(wi * StL::R + m) * TR + r];
      Dft<StL::R, false>::run(a);
#pragma unroll
      for (int m = 0; m < StL::R; ++m) {
        const size_t g = (size_t)(wi * StL::R + m) * M + i0 + r;
        if (SAVE_Z || FWD_ONLY) (FWD_ONLY ? A.out : A.zsave)[qoff + g] = a[m];
        if (ACC != 0) {
          const float2 z2 = A.z2[qoff + g];
          const float re = a[m].x * z2.x + a[m].y * z2.y, im = a[m].x * z2.y - a[m].y * z2.x;   // conj(Z) * Z2
          if (SMACC) gsm[(wi * StL::R + m) * TR + r] += re;
          else {
            float* pa = reinterpret_cast<float*>(A.pacc + toff + g);
            atomicAdd(pa, re);
            atomicAdd(pa + 1, im);
          }
        }
        if (!FWD_ONLY) {
          if (MUL == 0) a[m] = cscale(a[m], A.ctab[toff + g]);
          else {
            const float2 kk = A.ktab[toff + g];
            a[m] = cmul(a[m], make_float2(kk.x, MUL == 2 ? -kk.y : kk.y));
          }
        }
      }
      if (!FWD_ONLY) {
        Dft<StL::R, true>::run(a);
#pragma unroll
        for (int m = 0; m < StL::R; ++m) tile[(wi * StL::R + m) * TR + r] = a[m];
      }
    }
  }
}
// gacc += the block's shared-memory partial sums of G (same thread -> element ownership as dim2_fused_stage)
template <int LN, int NT, int TR>
ADMMTV_DI void dim2_fused_flush(const float* gsm, const Dim2Args& A, size_t toff, int i0, int tid) {
  constexpr int N = dim_len(LN), NS = plan_stages(N, kP2), RP = TR / 2;
  using StL = Stage<N, NS - 1, kP2>;
  const int M = A.M;
  if constexpr (StL::R <= 16) {
    for (int item = tid; item < RP * StL::ITEMS; item += NT) {
      const int rp = item % RP, wi = item / RP;
#pragma unroll
      for (int m = 0; m < StL::R; ++m) {
        const size_t g = (size_t)(wi * StL::R + m) * M + i0 + 2 * rp;
        const float2 gv = *reinterpret_cast<const float2*>(gsm + (wi * StL::R + m) * TR + 2 * rp);
        atomicAdd(A.gacc + toff + g, (double)gv.x);
        atomicAdd(A.gacc + toff + g + 1, (double)gv.y);
      }
    }
  } else {
    for (int item = tid; item < TR * StL::ITEMS; item += NT) {
      const int r = item % TR, wi = item / TR;
#pragma unroll
      for (int m = 0; m < StL::R; ++m)
        atomicAdd(A.gacc + toff + (size_t)(wi * StL::R + m) * M + i0 + r, (double)gsm[(wi * StL::R + m) * TR + r]);
    }
  }
}

template <int LN, int MUL, bool SAVE_Z, int ACC, bool FWD_ONLY>
__global__ void __launch_bounds__(Dim2Cfg<LN>::NT, Dim2Cfg<LN>::MINB) k_dim2(Dim2Args A) {
  using Cfg = Dim2Cfg<LN>;
  constexpr int N = Cfg::N, TR = Cfg::TR, NT = Cfg::NT, RP = TR / 2, NS = plan_stages(N, kP2);
  // G = sum over pairs of Re(conj(Z) Z2) is accumulated in shared memory across this block's
  // pairs (each element is owned by one thread) and flushed with one atomic per element at the
  // end, instead of one global atomic per element per pair.
  constexpr bool SMACC = ACC == 1;
  ADMMTV_DYN_SMEM(float2, tile);  // [N][TR] (+ float [N][TR] accumulator when SMACC)
  const int tid = threadIdx.x, M = A.M;
  const int i0 = blockIdx.x * TR;
  float* gsm = reinterpret_cast<float*>(tile + N * TR);

  for (int q = blockIdx.y; q < A.Q; q += gridDim.y) {
    const size_t qoff = (size_t)q * N * M;
    const size_t toff = (size_t)(q / A.Qg) * A.tab_stride;  // this group's tables / accumulators
    const float2* src = A.in + qoff + i0;
    const bool first_of_run = q == (int)blockIdx.y || (q - (int)gridDim.y) / A.Qg != q / A.Qg;
    const bool last_of_run = q + (int)gridDim.y >= A.Q || (q + (int)gridDim.y) / A.Qg != q / A.Qg;
    if (SMACC && first_of_run) {
      for (int t = tid; t < N * TR; t += NT) gsm[t] = 0.f;  // ordered before its first use by the barriers of the forward passes
    }
    constexpr int PFN = LN == 11 ? ADMMTV_PF_NEXT2_11 : ADMMTV_PF_NEXT;
    if (PFN > 0 && gridDim.y == (unsigned)A.Q) {
      // the input tile of the block that takes over this SM slot next
      const long nid = (long)blockIdx.y * gridDim.x + blockIdx.x + PFN;
      if (nid < (long)gridDim.x * gridDim.y) {
        const float2* nxt = A.in + (size_t)(nid / gridDim.x) * N * M + (size_t)(nid % gridDim.x) * TR;
        for (int col = tid; col < N; col += NT) l2_prefetch_line(nxt + (size_t)col * M);
      }
    }
    if (ACC != 0) {
      // the second spectrum is consumed in the fused stage: start pulling it into L2 now
      for (int col = tid; col < N; col += NT) l2_prefetch_line(A.z2 + qoff + i0 + (size_t)col * M);
    }
    if (q + (int)gridDim.y < A.Q) {
      // this block's next tile -> L2 (TR*8 bytes per column; one line covers 16 rows)
      const float2* nxt = A.in + (size_t)(q + gridDim.y) * N * M + i0;
      for (int col = tid; col < N; col += NT) l2_prefetch_line(nxt + (size_t)col * M);
    }

    // forward stage 0: global -> registers -> shared
    {
      using St = Stage<N, 0, kP2>;
      for (int item = tid; item < RP * St::ITEMS; item += NT) {
        const int rp = item % RP, wi = item / RP;
        float2 p[St::R];
        stage_twiddles<N, 0, false, kP2>(wi, A.twN, p);
        float2 a0[St::R], a1[St::R];
#pragma unroll
        for (int m = 0; m < St::R; ++m) {
          const float4 v = *reinterpret_cast<const float4*>(src + (size_t)(wi + m * St::STRIDE) * M + 2 * rp);
          a0[m] = make_float2(v.x, v.y);
          a1[m] = make_float2(v.z, v.w);
        }
        stage_fwd<N, 0, kP2>(a0, p);
        stage_fwd<N, 0, kP2>(a1, p);
#pragma unroll
        for (int m = 0; m < St::R; ++m)
          *reinterpret_cast<float4*>(tile + d2col<LN, TR, Cfg::SWZ>(wi + m * St::STRIDE) * TR + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
      }
    }
    __syncthreads();
    dim2_fwd_mid<LN, 1>(tile, A.twN, tid);

    // last forward stage fused with the spectral multiply and the first inverse stage
    dim2_fused_stage<LN, MUL, SAVE_Z, ACC, FWD_ONLY, NT, TR, Cfg::SWZ>(tile, gsm, A, qoff, toff, i0, tid);
    __syncthreads();
    if (FWD_ONLY) continue;
    dim2_inv_mid<LN, NS - 2>(tile, A.twN, tid);

    // inverse stage 0: shared -> registers -> global
    {
      using St = Stage<N, 0, kP2>;
      float2* dst = A.out + qoff + i0;
      for (int item = tid; item < RP * St::ITEMS; item += NT) {
        const int rp = item % RP, wi = item / RP;
        float2 p[St::R];
        stage_twiddles<N, 0, true, kP2>(wi, A.twN, p);
        float2 a0[St::R], a1[St::R];
#pragma unroll
        for (int m = 0; m < St::R; ++m) {
          const float4 v = *reinterpret_cast<const float4*>(tile + d2col<LN, TR, Cfg::SWZ>(wi + m * St::STRIDE) * TR + 2 * rp);
          a0[m] = make_float2(v.x, v.y);
          a1[m] = make_float2(v.z, v.w);
        }
        stage_inv<N, 0, kP2>(a0, p);
        stage_inv<N, 0, kP2>(a1, p);
#pragma unroll
        for (int m = 0; m < St::R; ++m)
          *reinterpret_cast<float4*>(dst + (size_t)(wi + m * St::STRIDE) * M + 2 * rp) = make_float4(a0[m].x, a0[m].y, a1[m].x, a1[m].y);
      }
    }
    if (SMACC && last_of_run) dim2_fused_flush<LN, NT, TR>(gsm, A, toff, i0, tid);
    __syncthreads();  // the tile is rewritten by the next pair's first pass
  }
}

}  // namespace admmtv
