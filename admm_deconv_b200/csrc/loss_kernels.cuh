// loss_kernels.cuh -- GMSD and SSIM, forward and hand-written backward.
//
// Replaces the cuDNN grouped convolutions + CUDA.jl broadcasts + Zygote tape of
//   /root/reference/src/metrics/gmsd.jl:13-27 (+ iqa_utils.jl:24-55) and src/metrics/ssim.jl:84-124
// by one kernel per direction: every intermediate map lives in registers (GMSD: warp strips marching along dim 2,
// neighbours by SHFL) or shared memory (SSIM: 32 x 32 output tiles, register-blocked separable windows), and the
// only HBM traffic is x, y in and xbar out (GMSD: 8 B/pixel forward, 12 B/pixel backward; SSIM additionally keeps
// three derivative maps, 12 B/output pixel, between forward and backward instead of redoing five 11x11 windows).
// Arrays are Julia (M,N,C,B) column-major: plane s = c + C*b, element (i,j) at i + M*j.
#pragma once

#include "compat.cuh"
#include "launch_macros.cuh"
#include "reduce.cuh"

namespace admmtv {

ADMMTV_DI int wrapi(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// ------------------------------------------------------------------------------------------
// GMSD
// ------------------------------------------------------------------------------------------
struct GmsdArgs {
  const float* x;
  const float* y;
  int M, N, C, B;
  int tiles_i, tiles_j;
  float t, alpha;
  double* acc;           // [2B]  sum (g-1), sum (g-1)^2 per image
  double* stats;         // [2B]  mean g, sqrt(score) per image
  const float* lossbar;  // backward
  float* out;            // forward: loss ; backward: xbar
};

// Sobel/8 gradients (iqa_utils.jl:15-20,46-47; NNlib conv = true convolution of the circularly padded image):
//   gx[i,j] = sum_dj w(dj) (x[i+1,j+dj] - x[i-1,j+dj]),  gy[i,j] = sum_di w(di) (x[i+di,j+1] - x[i+di,j-1]),  w = (1,2,1)/8
// evaluated from the separable column parts cs = x[i-1] + 2 x[i] + x[i+1], cd = x[i+1] - x[i-1] (gs_col below).
// The loss kernels are instruction-bound, so the per-pixel square root and divisions use the 2-ulp hardware approximations
// (MUFU.RSQ / MUFU.RCP); the error is far inside the loss tolerances (tests/test_gpu_losses.py).
ADMMTV_DI float gradmag(float gx, float gy) {   // iqa_utils.jl:53-55
  const float s = gx * gx + gy * gy + 1e-16f;
  return s * rsqrtf(s);
}

// per-image mean / deviation (gmsd.jl:21-24) and the batch mean (reduction = mean, :26)
__global__ void __launch_bounds__(128) k_gmsd_finalize(GmsdArgs A) {
  const double n = (double)A.M * A.N * A.C;
  double tot = 0.0;
  for (int b = threadIdx.x; b < A.B; b += 128) {
    const double m = A.acc[2 * b] / n;
    double var = A.acc[2 * b + 1] / n - m * m;
    if (var < 0.0) var = 0.0;
    const double sc = sqrt(var);
    A.stats[2 * b] = 1.0 + m;
    A.stats[2 * b + 1] = sc;
    tot += sc;
  }
  tot = block_sum(tot);
  if (threadIdx.x == 0) A.out[0] = (float)(tot / A.B);
}

// ------------------------------------------------------------------------------------------
// GMSD kernels, streaming form: no shared memory, no block barrier (round 1's 64 x 32 shared-memory tile kernels ran at
// 0.29 + 0.48 ms on 64 x 512^2 x 3; these at 0.18 + 0.29 ms, issue-bound at 75 % by the arithmetic of the metric itself).  A warp owns a strip of rows (one row per lane, the outer
// lanes are the circular halo) and marches along dim 2: each step is one coalesced load of a column of x and of y, the
// row neighbours come from the adjacent lanes (SHFL), the column neighbours from a three-column sliding window in
// registers.  The backward carries a second window (the adjoint stencil of the two cotangent fields) two columns behind.
// ------------------------------------------------------------------------------------------
#ifndef ADMMTV_GS_UNROLL_F
#define ADMMTV_GS_UNROLL_F 4
#endif
#ifndef ADMMTV_GS_UNROLL_B
#define ADMMTV_GS_UNROLL_B 2
#endif
#ifndef ADMMTV_GS_CW
#define ADMMTV_GS_CW 64
#endif
#ifndef ADMMTV_GS_NT
#define ADMMTV_GS_NT 256
#endif
constexpr int GS_NT = ADMMTV_GS_NT, GS_CW = ADMMTV_GS_CW;        // threads per block, columns per strip
constexpr int GS_UF = ADMMTV_GS_UNROLL_F, GS_UB = ADMMTV_GS_UNROLL_B;
constexpr int GS_RF = 30, GS_RB = 28;         // useful rows per warp: forward (1-row halo), backward (2-row halo)

// separable Sobel parts of one column at this lane's row: cs = up + 2 mid + down, cd = down - up
ADMMTV_DI void gs_col(float v, int lane, float& cs, float& cd) {
  const float up = __shfl_sync(0xffffffffu, v, (lane + 31) & 31), dn = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
  cs = up + 2.f * v + dn;
  cd = dn - up;
}

__global__ void __launch_bounds__(GS_NT) k_gmsd_fwd_s(GmsdArgs A) {
  const int lane = threadIdx.x & 31, strips = A.tiles_i * A.tiles_j;
  const long long wid = (long long)blockIdx.x * (GS_NT / 32) + (threadIdx.x >> 5);
  if (wid >= (long long)strips * A.C * A.B) return;   // warp-uniform
  const int s = (int)(wid / strips), st = (int)(wid % strips);
  const int r0 = (st % A.tiles_i) * GS_RF, c0 = (st / A.tiles_i) * GS_CW;
  const int row = r0 - 1 + lane;
  const bool row_ok = lane >= 1 && lane <= GS_RF && row < A.M;
  const size_t plane = (size_t)A.M * A.N;
  const float* xp = A.x + (size_t)s * plane + wrapi(row, A.M);
  const float* yp = A.y + (size_t)s * plane + wrapi(row, A.M);
  int gj = wrapi(c0 - 1, A.N);
  auto next = [&](int j) { return j + 1 >= A.N ? j + 1 - A.N : j + 1; };
  float xs0, xd0, xs1, xd1, ys0, yd0, ys1, yd1;
  gs_col(xp[(size_t)gj * A.M], lane, xs0, xd0); gs_col(yp[(size_t)gj * A.M], lane, ys0, yd0);
  gj = next(gj);
  gs_col(xp[(size_t)gj * A.M], lane, xs1, xd1); gs_col(yp[(size_t)gj * A.M], lane, ys1, yd1);
  float f1 = 0.f, f2 = 0.f;
  double d1 = 0.0, d2 = 0.0;
#pragma unroll GS_UF
  for (int k = 0; k < GS_CW; ++k) {
    gj = next(gj);
    float xs2, xd2, ys2, yd2;
    gs_col(xp[(size_t)gj * A.M], lane, xs2, xd2);
    gs_col(yp[(size_t)gj * A.M], lane, ys2, yd2);
    if (row_ok && c0 + k < A.N) {
      const float mx = gradmag((xd0 + 2.f * xd1 + xd2) * 0.125f, (xs2 - xs0) * 0.125f);
      const float my = gradmag((yd0 + 2.f * yd1 + yd2) * 0.125f, (ys2 - ys0) * 0.125f);
      const float mm = mx * my;
      const float g = __fdividef(2.f * mm - A.alpha * mm + A.t, mx * mx + my * my - A.alpha * mm + A.t);  // gmsd.jl:5-10
      const float d = g - 1.f;  // sums of (g-1): well conditioned when x ~ y
      f1 += d;
      f2 += d * d;
    }
    if ((k & 7) == 7) {   // fp32 over 8 pixels, fp64 beyond
      d1 += (double)f1; d2 += (double)f2;
      f1 = 0.f; f2 = 0.f;
    }
    xs0 = xs1; xd0 = xd1; xs1 = xs2; xd1 = xd2;
    ys0 = ys1; yd0 = yd1; ys1 = ys2; yd1 = yd2;
  }
  d1 = warp_sum(d1);
  d2 = warp_sum(d2);
  if (lane == 0) {
    const int b = s / A.C;
    atomicAdd(A.acc + 2 * b, d1);
    atomicAdd(A.acc + 2 * b + 1, d2);
  }
}

__global__ void __launch_bounds__(GS_NT) k_gmsd_bwd_s(GmsdArgs A) {
  const int lane = threadIdx.x & 31, strips = A.tiles_i * A.tiles_j;
  const long long wid = (long long)blockIdx.x * (GS_NT / 32) + (threadIdx.x >> 5);
  if (wid >= (long long)strips * A.C * A.B) return;   // warp-uniform
  const int s = (int)(wid / strips), st = (int)(wid % strips);
  const int r0 = (st % A.tiles_i) * GS_RB, c0 = (st / A.tiles_i) * GS_CW;
  const int row = r0 - 2 + lane;
  const bool row_ok = lane >= 2 && lane <= GS_RB + 1 && row < A.M;
  const size_t plane = (size_t)A.M * A.N;
  const float* xp = A.x + (size_t)s * plane + wrapi(row, A.M);
  const float* yp = A.y + (size_t)s * plane + wrapi(row, A.M);
  float* xb = A.out + (size_t)s * plane + (row_ok ? row : 0);
  const int b = s / A.C;
  const float mean = (float)A.stats[2 * b];
  // d loss / d g[p] = lossbar (g[p] - mean_b) / (B n sqrt(score_b))
  const float scale = (float)((double)A.lossbar[0] / ((double)A.B * ((double)A.M * A.N * A.C) * A.stats[2 * b + 1]));
  int gj = wrapi(c0 - 2, A.N);
  auto next = [&](int j) { return j + 1 >= A.N ? j + 1 - A.N : j + 1; };
  float xs0, xd0, xs1, xd1, ys0, yd0, ys1, yd1;
  gs_col(xp[(size_t)gj * A.M], lane, xs0, xd0); gs_col(yp[(size_t)gj * A.M], lane, ys0, yd0);
  gj = next(gj);
  gs_col(xp[(size_t)gj * A.M], lane, xs1, xd1); gs_col(yp[(size_t)gj * A.M], lane, ys1, yd1);
  float A0 = 0.f, B0 = 0.f, A1 = 0.f, B1 = 0.f;   // adjoint-stencil parts of the two previous cotangent columns
#pragma unroll GS_UB
  for (int t = 2; t < GS_CW + 4; ++t) {
    gj = next(gj);                       // image column c0 - 2 + t
    float xs2, xd2, ys2, yd2;
    gs_col(xp[(size_t)gj * A.M], lane, xs2, xd2);
    gs_col(yp[(size_t)gj * A.M], lane, ys2, yd2);
    // cotangents of the two gradient fields at column c0 + t - 3 (this lane's row)
    const float gx = (xd0 + 2.f * xd1 + xd2) * 0.125f, gy = (xs2 - xs0) * 0.125f;
    const float mx = gradmag(gx, gy);
    const float my = gradmag((yd0 + 2.f * yd1 + yd2) * 0.125f, (ys2 - ys0) * 0.125f);
    const float mm = mx * my;
    const float den = mx * mx + my * my - A.alpha * mm + A.t;
    const float rden = __fdividef(1.f, den);
    const float g = (2.f * mm - A.alpha * mm + A.t) * rden;
    const float dgdmx = ((2.f - A.alpha) * my - g * (2.f * mx - A.alpha * my)) * rden;
    const float c = __fdividef(scale * (g - mean) * dgdmx, mx);
    const float p1 = c * gx, p2 = c * gy;
    // adjoint of the two circular stencils: row parts of this column ...
    const float p1u = __shfl_sync(0xffffffffu, p1, (lane + 31) & 31), p1d = __shfl_sync(0xffffffffu, p1, (lane + 1) & 31);
    const float p2u = __shfl_sync(0xffffffffu, p2, (lane + 31) & 31), p2d = __shfl_sync(0xffffffffu, p2, (lane + 1) & 31);
    const float A2 = p1u - p1d, B2 = p2u + 2.f * p2 + p2d;
    // ... column parts over the three-column window: xbar at column c0 + t - 4
    const int oc = c0 + t - 4;
    if (t >= 4 && row_ok && oc < A.N) xb[(size_t)oc * A.M] = ((A0 + 2.f * A1 + A2) + (B0 - B2)) * 0.125f;
    A0 = A1; B0 = B1; A1 = A2; B1 = B2;
    xs0 = xs1; xd0 = xd1; xs1 = xs2; xd1 = xd2;
    ys0 = ys1; yd0 = yd1; ys1 = ys2; yd1 = yd2;
  }
}

// ------------------------------------------------------------------------------------------
// SSIM
// ------------------------------------------------------------------------------------------
constexpr int SS_T = 32, SS_NT = 256, SS_LMAX = 11, SS_IN = SS_T + SS_LMAX - 1;

struct SsimArgs {
  const float* x;
  const float* y;
  int M, N, C, B, L, Mo, No;
  int tiles_i, tiles_j;
  float f[SS_LMAX];  // flipped taps: NNlib conv is a true convolution (ssim.jl:112-119)
  float C1, C2;
  double* acc;        // [1] sum of the ssim map
  float* maps;        // [3][planes][No][Mo] dS/dmu_x, dS/dE[x^2], dS/dE[xy]  (with_grad)
  int with_grad, as_loss;
  const float* outbar;
  float* out;         // forward: ssim or 1-ssim ; backward: xbar
};

// One output pixel of the SSIM map from the five window means (ssim.jl:114-121); with_grad: its three derivative maps.
ADMMTV_DI float ssim_point(const SsimArgs& A, float mx, float my, float exx, float eyy, float exy, size_t o, size_t mapstride) {
  const float mxy = mx * my, mx2 = mx * mx, my2 = my * my;
  const float sx2 = exx - mx2, sy2 = eyy - my2, sxy = exy - mxy;                   // ssim.jl:117-119
  const float A1 = 2.f * mxy + A.C1, A2 = 2.f * sxy + A.C2, B1 = mx2 + my2 + A.C1, B2 = sx2 + sy2 + A.C2;
  // the kernels are instruction-bound: the reciprocals are the 1-ulp hardware approximation (MUFU.RCP), far inside the tolerances
  const float r1 = __fdividef(1.f, B1), r2 = __fdividef(1.f, B2), ib = r1 * r2;
  const float S = A1 * A2 * ib;                                                    // ssim.jl:121
  if (A.with_grad) {
    A.maps[o] = 2.f * my * (A2 - A1) * ib - 2.f * mx * S * (r1 - r2);               // dS/dmu_x
    A.maps[mapstride + o] = -S * r2;                                               // dS/dE[x^2]
    A.maps[2 * mapstride + o] = 2.f * A1 * ib;                                     // dS/dE[xy]
  }
  return S;
}

// Equal separable taps with a RUN-TIME count (LC = 0: A.L taps); the 11- and 5-tap windows of the reference take k_ssim_fwd4 / k_ssim_bwd4 below.
template <int LC>
__global__ void __launch_bounds__(SS_NT) k_ssim_fwd(SsimArgs A) {
  __shared__ float xs[SS_IN * SS_IN], ys[SS_IN * SS_IN];
  __shared__ float P[5][SS_IN * SS_T];
  const int tid = threadIdx.x, tiles = A.tiles_i * A.tiles_j, L = LC > 0 ? LC : A.L;
  const int s = blockIdx.x / tiles, tl = blockIdx.x % tiles;
  const int i0 = (tl % A.tiles_i) * SS_T, j0 = (tl / A.tiles_i) * SS_T;
  const size_t plane = (size_t)A.M * A.N;
  const float* xp = A.x + (size_t)s * plane;
  const float* yp = A.y + (size_t)s * plane;
  const int ext = SS_T + L - 1;
  for (int e = tid; e < ext * ext; e += SS_NT) {
    const int li = e % ext, lj = e / ext;
    const int gi = i0 + li, gj = j0 + lj;
    const bool ok = gi < A.M && gj < A.N;
    xs[lj * SS_IN + li] = ok ? xp[(size_t)gj * A.M + gi] : 0.f;
    ys[lj * SS_IN + li] = ok ? yp[(size_t)gj * A.M + gi] : 0.f;
  }
  __syncthreads();
  // window along dim 1 for x, y, x^2, y^2, xy
  for (int e = tid; e < SS_T * ext; e += SS_NT) {
    const int li = e % SS_T, lj = e / SS_T;
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int a = 0; a < L; ++a) {
      const float w = A.f[a], xv = xs[lj * SS_IN + li + a], yv = ys[lj * SS_IN + li + a];
      sx += w * xv; sy += w * yv; sxx += w * xv * xv; syy += w * yv * yv; sxy += w * xv * yv;
    }
    P[0][e] = sx; P[1][e] = sy; P[2][e] = sxx; P[3][e] = syy; P[4][e] = sxy;
  }
  __syncthreads();
  double tot = 0.0;
  const size_t oplane = (size_t)A.Mo * A.No, nplanes = (size_t)A.C * A.B;
  for (int e = tid; e < SS_T * SS_T; e += SS_NT) {
    const int li = e % SS_T, lj = e / SS_T;
    if (i0 + li < A.Mo && j0 + lj < A.No) {
      float mx = 0.f, my = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
  #pragma unroll
    for (int a = 0; a < L; ++a) {
        const float w = A.f[a];
        const int o = (lj + a) * SS_T + li;
        mx += w * P[0][o]; my += w * P[1][o]; exx += w * P[2][o]; eyy += w * P[3][o]; exy += w * P[4][o];
      }
      tot += (double)ssim_point(A, mx, my, exx, eyy, exy, (size_t)s * oplane + (size_t)(j0 + lj) * A.Mo + i0 + li, nplanes * oplane);
    }
  }
  tot = block_sum(tot);
  if (tid == 0) atomicAdd(A.acc, tot);
}

__global__ void k_ssim_finalize(SsimArgs A) {
  // mean over (1,2,3) then over the batch: equal-sized images => the global mean (ssim.jl:122-123)
  const double m = A.acc[0] / ((double)A.Mo * A.No * A.C * A.B);
  A.out[0] = (float)(A.as_loss ? 1.0 - m : m);
}

template <int LC>
__global__ void __launch_bounds__(SS_NT) k_ssim_bwd(SsimArgs A) {
  __shared__ float ms[3][SS_IN * SS_IN];
  __shared__ float T[3][SS_IN * SS_T];
  const int tid = threadIdx.x, tiles = A.tiles_i * A.tiles_j, L = LC > 0 ? LC : A.L;
  const int s = blockIdx.x / tiles, tl = blockIdx.x % tiles;
  const int i0 = (tl % A.tiles_i) * SS_T, j0 = (tl / A.tiles_i) * SS_T;
  const size_t plane = (size_t)A.M * A.N, oplane = (size_t)A.Mo * A.No, nplanes = (size_t)A.C * A.B;
  const int ext = SS_T + L - 1;
  // derivative maps over output positions [i0-L+1, i0+T) x [j0-L+1, j0+T), zero outside the valid region
  for (int e = tid; e < ext * ext; e += SS_NT) {
    const int li = e % ext, lj = e / ext;
    const int pi = i0 - (L - 1) + li, pj = j0 - (L - 1) + lj;
    const bool ok = pi >= 0 && pj >= 0 && pi < A.Mo && pj < A.No;
    const size_t o = (size_t)s * oplane + (size_t)(ok ? pj : 0) * A.Mo + (ok ? pi : 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) ms[k][lj * SS_IN + li] = ok ? A.maps[k * nplanes * oplane + o] : 0.f;
  }
  __syncthreads();
  // transposed window along dim 1: t[qi, pj] = sum_a f[a] map[qi - a, pj]
  for (int e = tid; e < SS_T * ext; e += SS_NT) {
    const int li = e % SS_T, lj = e / SS_T;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int a = 0; a < L; ++a) {
      const float w = A.f[a];
      const int o = lj * SS_IN + li + (L - 1) - a;
      r0 += w * ms[0][o]; r1 += w * ms[1][o]; r2 += w * ms[2][o];
    }
    T[0][e] = r0; T[1][e] = r1; T[2][e] = r2;
  }
  __syncthreads();
  const float scale = (float)((double)A.outbar[0] * (A.as_loss ? -1.0 : 1.0) / ((double)A.Mo * A.No * A.C * A.B));
  const float* xp = A.x + (size_t)s * plane;
  const float* yp = A.y + (size_t)s * plane;
  float* xb = A.out + (size_t)s * plane;
  for (int e = tid; e < SS_T * SS_T; e += SS_NT) {
    const int li = e % SS_T, lj = e / SS_T;
    const int gi = i0 + li, gj = j0 + lj;
    if (gi < A.M && gj < A.N) {
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
  #pragma unroll
    for (int a = 0; a < L; ++a) {
        const float w = A.f[a];
        const int o = (lj + (L - 1) - a) * SS_T + li;
        r0 += w * T[0][o]; r1 += w * T[1][o]; r2 += w * T[2][o];
      }
      const size_t g = (size_t)gj * A.M + gi;
      xb[g] = scale * (r0 + 2.f * xp[g] * r1 + yp[g] * r2);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Register-blocked SSIM kernels for a compile-time tap count (11 = the default Gaussian, 5 = ssim_loss_fast).
// k_ssim_fwd / k_ssim_bwd above spend most of their time on shared-memory loads: L per window sum, 84 per output pixel in the
// forward.  Here every thread produces FOUR adjacent outputs of a window pass from the L + 3 inputs they share: along dim 1
// (contiguous in shared memory) the inputs come in as aligned LDS.128, along dim 2 as L + 3 conflict-free scalar loads, and the
// 4 x L products are register FMAs with the taps as constant-bank operands -- 2.6x fewer shared-memory wavefronts per pixel.
// The row stride SS_LD = 44 floats keeps every row 16-byte aligned.
// ------------------------------------------------------------------------------------------
constexpr int SS_LD = 44;

// Packed dual-fp32 arithmetic (PTX fma / mul .f32x2 -> SASS FFMA2 / FMUL2, Blackwell): the two images go through identical
// window sums, so (x, y) and (x^2, y^2) travel as pairs and one issue slot serves both.  Scalar fallback for the emulation.
#ifndef ADMMTV_LOSS_F32X2
#define ADMMTV_LOSS_F32X2 1
#endif
#if ADMMTV_LOSS_F32X2 && !defined(ADMMTV_EMU)
ADMMTV_DI float2 ffma2(float2 a, float2 b, float2 c) {   // a * b + c, componentwise
  unsigned long long ra, rb, rc, rd;
  float2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
  return r;
}
ADMMTV_DI float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  float2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
  return r;
}
#else
ADMMTV_DI float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
ADMMTV_DI float2 fmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif

// out[k] = sum_a f[a] v[k + a] (REV: v[k + L-1-a]),  k = 0..3
template <int L, bool REV>
ADMMTV_DI void win4(const float* v, const float* f, float* out) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float r = 0.f;
#pragma unroll
    for (int a = 0; a < L; ++a) r += f[a] * v[k + (REV ? L - 1 - a : a)];
    out[k] = r;
  }
}
// the same on pairs; w[a] = (f[a], f[a])
template <int L, bool REV>
ADMMTV_DI void win4x2(const float2* v, const float2* w, float2* out) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < L; ++a) r = ffma2(w[a], v[k + (REV ? L - 1 - a : a)], r);
    out[k] = r;
  }
}
// v[0 .. 2 NV) = row[0 .. 2 NV) of pairs, row 16-byte aligned
template <int NV>
ADMMTV_DI void load_row2(const float2* row, float2* v) {
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(row + 2 * q);
    v[2 * q] = make_float2(t.x, t.y);
    v[2 * q + 1] = make_float2(t.z, t.w);
  }
}
// v[0 .. 4 NV) = row[0 .. 4 NV), row 16-byte aligned
template <int NV>
ADMMTV_DI void load_row4(const float* row, float* v) {
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(row + 4 * q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

template <int L>
__global__ void __launch_bounds__(SS_NT) k_ssim_fwd4(SsimArgs A) {
  constexpr int ext = SS_T + L - 1, NE = 4 * ((L + 3 + 3) / 4);   // elements a thread pulls in per row: L + 3, rounded up to 4
  static_assert(NE + SS_T - 4 <= SS_LD && ext <= SS_IN, "a thread's aligned run stays inside its row");
  __shared__ __align__(16) float2 xy[SS_IN * SS_LD];       // (x, y) pairs of the input tile
  __shared__ __align__(16) float2 P01[SS_IN * SS_T];       // dim-1 window of (x, y)
  __shared__ __align__(16) float2 P23[SS_IN * SS_T];       //              of (x^2, y^2)
  __shared__ __align__(16) float P4[SS_IN * SS_T];         //              of x y
  const int tid = threadIdx.x, tiles = A.tiles_i * A.tiles_j;
  const int s = blockIdx.x / tiles, tl = blockIdx.x % tiles;
  const int i0 = (tl % A.tiles_i) * SS_T, j0 = (tl / A.tiles_i) * SS_T;
  const size_t plane = (size_t)A.M * A.N;
  const float* xp = A.x + (size_t)s * plane;
  const float* yp = A.y + (size_t)s * plane;
  float2 w[L];
#pragma unroll
  for (int a = 0; a < L; ++a) w[a] = make_float2(A.f[a], A.f[a]);
  {   // whole rows incl. the alignment padding (zeros): thread = (element lc of a row, every 4th row from lr)
    const int lc = tid % 64, lr = tid / 64;
    const bool col_ok = lc < ext && i0 + lc < A.M;
    if (lc < SS_LD) {
      const float* xq = xp + (size_t)(j0 + lr) * A.M + i0 + lc;
      const float* yq = yp + (size_t)(j0 + lr) * A.M + i0 + lc;
#pragma unroll 4
      for (int lj = lr; lj < ext; lj += 4, xq += 4 * (size_t)A.M, yq += 4 * (size_t)A.M) {
        const bool ok = col_ok && j0 + lj < A.N;
        xy[lj * SS_LD + lc] = ok ? make_float2(*xq, *yq) : make_float2(0.f, 0.f);
      }
    }
  }
  __syncthreads();
  // window along dim 1 for (x, y), (x^2, y^2), xy: item = (4 adjacent rows li0.., column lj)
  for (int e = tid; e < (SS_T / 4) * ext; e += SS_NT) {
    const int li0 = 4 * (e % (SS_T / 4)), lj = e / (SS_T / 4);
    float2 v[NE], q[NE], o[4];
    float pv[NE], os[4];
    load_row2<NE / 2>(xy + lj * SS_LD + li0, v);
    win4x2<L, false>(v, w, o);
    float2* d01 = P01 + lj * SS_T + li0;
    *reinterpret_cast<float4*>(d01) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
    *reinterpret_cast<float4*>(d01 + 2) = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
#pragma unroll
    for (int k = 0; k < L + 3; ++k) {
      q[k] = fmul2(v[k], v[k]);
      pv[k] = v[k].x * v[k].y;
    }
    win4x2<L, false>(q, w, o);
    float2* d23 = P23 + lj * SS_T + li0;
    *reinterpret_cast<float4*>(d23) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
    *reinterpret_cast<float4*>(d23 + 2) = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
    win4<L, false>(pv, A.f, os);
    *reinterpret_cast<float4*>(P4 + lj * SS_T + li0) = make_float4(os[0], os[1], os[2], os[3]);
  }
  __syncthreads();
  // window along dim 2: thread = (row li, 4 adjacent columns lj0..)
  static_assert(SS_NT == SS_T * (SS_T / 4), "one thread per (row, 4-column group)");
  const int li = tid % SS_T, lj0 = 4 * (tid / SS_T);
  float2 m01[4], m23[4];
  float m4[4];
  {
    float2 v[L + 3];
#pragma unroll
    for (int k = 0; k < L + 3; ++k) v[k] = P01[(lj0 + k) * SS_T + li];
    win4x2<L, false>(v, w, m01);
#pragma unroll
    for (int k = 0; k < L + 3; ++k) v[k] = P23[(lj0 + k) * SS_T + li];
    win4x2<L, false>(v, w, m23);
    float u[L + 3];
#pragma unroll
    for (int k = 0; k < L + 3; ++k) u[k] = P4[(lj0 + k) * SS_T + li];
    win4<L, false>(u, A.f, m4);
  }
  double tot = 0.0;
  const size_t oplane = (size_t)A.Mo * A.No, nplanes = (size_t)A.C * A.B;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i0 + li < A.Mo && j0 + lj0 + k < A.No)
      tot += (double)ssim_point(A, m01[k].x, m01[k].y, m23[k].x, m23[k].y, m4[k],
                                (size_t)s * oplane + (size_t)(j0 + lj0 + k) * A.Mo + i0 + li, nplanes * oplane);
  tot = block_sum(tot);
  if (tid == 0) atomicAdd(A.acc, tot);
}

// (the packed pairs lose in the backward: 0.576 -> 0.606 ms, three maps from three planes; it keeps scalar windows)
template <int L>
__global__ void __launch_bounds__(SS_NT) k_ssim_bwd4(SsimArgs A) {
  constexpr int ext = SS_T + L - 1, NV = (L + 3 + 3) / 4;
  __shared__ __align__(16) float ms[3][SS_IN * SS_LD];
  __shared__ __align__(16) float T[3][SS_IN * SS_T];
  const int tid = threadIdx.x, tiles = A.tiles_i * A.tiles_j;
  const int s = blockIdx.x / tiles, tl = blockIdx.x % tiles;
  const int i0 = (tl % A.tiles_i) * SS_T, j0 = (tl / A.tiles_i) * SS_T;
  const size_t plane = (size_t)A.M * A.N, oplane = (size_t)A.Mo * A.No, nplanes = (size_t)A.C * A.B;
  // derivative maps over output positions [i0-L+1, i0+T) x [j0-L+1, j0+T), zero outside the valid region
  {
    const int lc = tid % 64, lr = tid / 64;
    const int pi = i0 - (L - 1) + lc;
    const bool col_ok = lc < ext && pi >= 0 && pi < A.Mo;
    if (lc < SS_LD) {
      const float* mq = A.maps + (size_t)s * oplane + (col_ok ? pi : 0);
#pragma unroll 2
      for (int lj = lr; lj < ext; lj += 4) {
        const int pj = j0 - (L - 1) + lj;
        const bool ok = col_ok && pj >= 0 && pj < A.No;
        const float* m0 = mq + (size_t)(ok ? pj : 0) * A.Mo;
#pragma unroll
        for (int k = 0; k < 3; ++k) ms[k][lj * SS_LD + lc] = ok ? m0[k * nplanes * oplane] : 0.f;
      }
    }
  }
  __syncthreads();
  // transposed window along dim 1: t[qi, pj] = sum_a f[a] map[qi - a, pj]  (tile index li + L-1 - a)
  for (int e = tid; e < (SS_T / 4) * ext; e += SS_NT) {
    const int li0 = 4 * (e % (SS_T / 4)), lj = e / (SS_T / 4);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v[4 * NV], o[4];
      load_row4<NV>(ms[c] + lj * SS_LD + li0, v);
      win4<L, true>(v, A.f, o);
      *reinterpret_cast<float4*>(&T[c][lj * SS_T + li0]) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  const float scale = (float)((double)A.outbar[0] * (A.as_loss ? -1.0 : 1.0) / ((double)A.Mo * A.No * A.C * A.B));
  const float* xp = A.x + (size_t)s * plane;
  const float* yp = A.y + (size_t)s * plane;
  float* xb = A.out + (size_t)s * plane;
  const int li = tid % SS_T, lj0 = 4 * (tid / SS_T);
  float r[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[L + 3];
#pragma unroll
    for (int q = 0; q < L + 3; ++q) v[q] = T[c][(lj0 + q) * SS_T + li];
    win4<L, true>(v, A.f, r[c]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int gi = i0 + li, gj = j0 + lj0 + k;
    if (gi < A.M && gj < A.N) {
      const size_t g = (size_t)gj * A.M + gi;
      xb[g] = scale * (r[0][k] + 2.f * xp[g] * r[1][k] + yp[g] * r[2][k]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// SSIM with an ARBITRARY 2-D window (ssim.jl:84 `kernel_ref`: any (L1, L2) array, e.g. a non-separable or non-square one).
// The host factors the window into R <= min(L1, L2) separable terms  W[a,b] = sum_r u_r[a] v_r[b]  (SVD); convolution is linear
// in the window, so each of the five window means is the sum of R separable passes through the same shared-memory tile.
// ------------------------------------------------------------------------------------------
struct SsimGenArgs {
  SsimArgs a;                      // f[] unused; Mo = M - L1 + 1, No = N - L2 + 1
  int L1, L2, R;
  float fu[SS_LMAX * SS_LMAX];     // [R][SS_LMAX] flipped dim-1 factors
  float fv[SS_LMAX * SS_LMAX];     // [R][SS_LMAX] flipped dim-2 factors
};
constexpr int SS_PT = SS_T * SS_T / SS_NT;   // output pixels per thread

__global__ void __launch_bounds__(SS_NT) k_ssim_fwd_gen(SsimGenArgs G) {
  const SsimArgs& A = G.a;
  __shared__ float xs[SS_IN * SS_IN], ys[SS_IN * SS_IN];
  __shared__ float P[5][SS_IN * SS_T];
  const int tid = threadIdx.x, tiles = A.tiles_i * A.tiles_j, L1 = G.L1, L2 = G.L2;
  const int s = blockIdx.x / tiles, tl = blockIdx.x % tiles;
  const int i0 = (tl % A.tiles_i) * SS_T, j0 = (tl / A.tiles_i) * SS_T;
  const size_t plane = (size_t)A.M * A.N;
  const float* xp = A.x + (size_t)s * plane;
  const float* yp = A.y + (size_t)s * plane;
  const int ext_i = SS_T + L1 - 1, ext_j = SS_T + L2 - 1;
  for (int e = tid; e < ext_i * ext_j; e += SS_NT) {
    const int li = e % ext_i, lj = e / ext_i;
    const int gi = i0 + li, gj = j0 + lj;
    const bool ok = gi < A.M && gj < A.N;
    xs[lj * SS_IN + li] = ok ? xp[(size_t)gj * A.M + gi] : 0.f;
    ys[lj * SS_IN + li] = ok ? yp[(size_t)gj * A.M + gi] : 0.f;
  }
  float acc[SS_PT][5];
#pragma unroll
  for (int k = 0; k < SS_PT; ++k)
#pragma unroll
    for (int c = 0; c < 5; ++c) acc[k][c] = 0.f;
  for (int r = 0; r < G.R; ++r) {
    __syncthreads();   // the tile is loaded / the previous term's second pass has read P
    const float* fu = G.fu + r * SS_LMAX;
    const float* fv = G.fv + r * SS_LMAX;
    for (int e = tid; e < SS_T * ext_j; e += SS_NT) {
      const int li = e % SS_T, lj = e / SS_T;
      float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
      for (int a = 0; a < L1; ++a) {
        const float w = fu[a], xv = xs[lj * SS_IN + li + a], yv = ys[lj * SS_IN + li + a];
        sx += w * xv; sy += w * yv; sxx += w * xv * xv; syy += w * yv * yv; sxy += w * xv * yv;
      }
      P[0][e] = sx; P[1][e] = sy; P[2][e] = sxx; P[3][e] = syy; P[4][e] = sxy;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SS_PT; ++k) {
      const int e = tid + k * SS_NT, li = e % SS_T, lj = e / SS_T;
      for (int b = 0; b < L2; ++b) {
        const float w = fv[b];
        const int o = (lj + b) * SS_T + li;
        acc[k][0] += w * P[0][o]; acc[k][1] += w * P[1][o]; acc[k][2] += w * P[2][o]; acc[k][3] += w * P[3][o]; acc[k][4] += w * P[4][o];
      }
    }
  }
  double tot = 0.0;
  const size_t oplane = (size_t)A.Mo * A.No, nplanes = (size_t)A.C * A.B;
#pragma unroll
  for (int k = 0; k < SS_PT; ++k) {
    const int e = tid + k * SS_NT, li = e % SS_T, lj = e / SS_T;
    if (i0 + li < A.Mo && j0 + lj < A.No)
      tot += (double)ssim_point(A, acc[k][0], acc[k][1], acc[k][2], acc[k][3], acc[k][4],
                                (size_t)s * oplane + (size_t)(j0 + lj) * A.Mo + i0 + li, nplanes * oplane);
  }
  tot = block_sum(tot);
  if (tid == 0) atomicAdd(A.acc, tot);
}

__global__ void __launch_bounds__(SS_NT) k_ssim_bwd_gen(SsimGenArgs G) {
  const SsimArgs& A = G.a;
  __shared__ float ms[3][SS_IN * SS_IN];
  __shared__ float T[3][SS_IN * SS_T];
  const int tid = threadIdx.x, tiles = A.tiles_i * A.tiles_j, L1 = G.L1, L2 = G.L2;
  const int s = blockIdx.x / tiles, tl = blockIdx.x % tiles;
  const int i0 = (tl % A.tiles_i) * SS_T, j0 = (tl / A.tiles_i) * SS_T;
  const size_t plane = (size_t)A.M * A.N, oplane = (size_t)A.Mo * A.No, nplanes = (size_t)A.C * A.B;
  const int ext_i = SS_T + L1 - 1, ext_j = SS_T + L2 - 1;
  // derivative maps over output positions [i0-L1+1, i0+T) x [j0-L2+1, j0+T), zero outside the valid region
  for (int e = tid; e < ext_i * ext_j; e += SS_NT) {
    const int li = e % ext_i, lj = e / ext_i;
    const int pi = i0 - (L1 - 1) + li, pj = j0 - (L2 - 1) + lj;
    const bool ok = pi >= 0 && pj >= 0 && pi < A.Mo && pj < A.No;
    const size_t o = (size_t)s * oplane + (size_t)(ok ? pj : 0) * A.Mo + (ok ? pi : 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) ms[k][lj * SS_IN + li] = ok ? A.maps[k * nplanes * oplane + o] : 0.f;
  }
  float acc[SS_PT][3];
#pragma unroll
  for (int k = 0; k < SS_PT; ++k) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
  for (int r = 0; r < G.R; ++r) {
    __syncthreads();
    const float* fu = G.fu + r * SS_LMAX;
    const float* fv = G.fv + r * SS_LMAX;
    // transposed window along dim 1: t[qi, pj] = sum_a fu[a] map[qi - a, pj]
    for (int e = tid; e < SS_T * ext_j; e += SS_NT) {
      const int li = e % SS_T, lj = e / SS_T;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
      for (int a = 0; a < L1; ++a) {
        const float w = fu[a];
        const int o = lj * SS_IN + li + (L1 - 1) - a;
        r0 += w * ms[0][o]; r1 += w * ms[1][o]; r2 += w * ms[2][o];
      }
      T[0][e] = r0; T[1][e] = r1; T[2][e] = r2;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SS_PT; ++k) {
      const int e = tid + k * SS_NT, li = e % SS_T, lj = e / SS_T;
      for (int b = 0; b < L2; ++b) {
        const float w = fv[b];
        const int o = (lj + (L2 - 1) - b) * SS_T + li;
        acc[k][0] += w * T[0][o]; acc[k][1] += w * T[1][o]; acc[k][2] += w * T[2][o];
      }
    }
  }
  const float scale = (float)((double)A.outbar[0] * (A.as_loss ? -1.0 : 1.0) / ((double)A.Mo * A.No * A.C * A.B));
  const float* xp = A.x + (size_t)s * plane;
  const float* yp = A.y + (size_t)s * plane;
  float* xb = A.out + (size_t)s * plane;
#pragma unroll
  for (int k = 0; k < SS_PT; ++k) {
    const int e = tid + k * SS_NT, li = e % SS_T, lj = e / SS_T;
    const int gi = i0 + li, gj = j0 + lj;
    if (gi < A.M && gj < A.N) {
      const size_t g = (size_t)gj * A.M + gi;
      xb[g] = scale * (acc[k][0] + 2.f * xp[g] * acc[k][1] + yp[g] * acc[k][2]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// pad_symmetric (NNlib; ssim.jl:104-110 `crop = false`): the border values are mirrored INCLUDING the edge sample,
//   dst[i, j] = src[mir(i - lo1, M), mir(j - lo2, N)],   mir(t, n) = t < 0 ? -1 - t : (t >= n ? 2n - 1 - t : t)
// and its adjoint (the pullback Zygote derives): every source pixel collects the cotangents of its mirror images.
// ------------------------------------------------------------------------------------------
struct PadArgs {
  const float* src;
  float* dst;
  int M, N, planes;       // unpadded plane size
  int lo1, hi1, lo2, hi2;
};
ADMMTV_DI int mir_sym(int t, int n) { return t < 0 ? -1 - t : (t >= n ? 2 * n - 1 - t : t); }

__global__ void __launch_bounds__(256) k_pad_symmetric(PadArgs A) {
  const int Mp = A.M + A.lo1 + A.hi1, Np = A.N + A.lo2 + A.hi2;
  const size_t total = (size_t)Mp * Np * A.planes;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int i = (int)(e % Mp), j = (int)((e / Mp) % Np);
    const size_t s = e / ((size_t)Mp * Np);
    A.dst[e] = A.src[s * A.M * A.N + (size_t)mir_sym(j - A.lo2, A.N) * A.M + mir_sym(i - A.lo1, A.M)];
  }
}
// src = the PADDED cotangent, dst = the unpadded one (fully overwritten)
__global__ void __launch_bounds__(256) k_pad_symmetric_adj(PadArgs A) {
  const int Mp = A.M + A.lo1 + A.hi1, Np = A.N + A.lo2 + A.hi2;
  const size_t total = (size_t)A.M * A.N * A.planes;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int i = (int)(e % A.M), j = (int)((e / A.M) % A.N);
    const size_t s = e / ((size_t)A.M * A.N);
    const float* p = A.src + s * Mp * Np;
    // padded positions along each dimension that mirror onto (i, j): itself, the low pad, the high pad
    int pi[3], pj[3], ni = 0, nj = 0;
    pi[ni++] = A.lo1 + i;
    if (i < A.lo1) pi[ni++] = A.lo1 - 1 - i;
    if (A.M - 1 - i < A.hi1) pi[ni++] = A.lo1 + 2 * A.M - 1 - i;
    pj[nj++] = A.lo2 + j;
    if (j < A.lo2) pj[nj++] = A.lo2 - 1 - j;
    if (A.N - 1 - j < A.hi2) pj[nj++] = A.lo2 + 2 * A.N - 1 - j;
    float v = 0.f;
    for (int b = 0; b < nj; ++b)
      for (int a = 0; a < ni; ++a) v += p[(size_t)pj[b] * Mp + pi[a]];
    A.dst[e] = v;
  }
}

// ------------------------------------------------------------------------------------------
// batch assembly: N0f8 channel-interleaved crops -> fp32 (M,N,C,B)
//   replaces img2tensor (base_funcs.jl:29-35) + cat(dims=4) (datafeeder.jl:54-68) + the fp32 upload
// One block = one 32x32 pixel tile of one image, four channels at a time; the tile is transposed through shared
// memory so that both the byte reads (source's fastest pixel dimension) and the float writes (dim 1) coalesce.
// ------------------------------------------------------------------------------------------
constexpr int BA_T = 32, BA_NT = 256, BA_CC = 4;

struct BatchArgs {
  const uint8_t* src;
  float* dst;
  int M, N, C, B;
  int tiles_i, tiles_j;
  long long sc, si, sj, sb;
  const long long* offsets;   // optional per-image element offsets (device-resident dataset gather); else b * sb
};

__global__ void __launch_bounds__(BA_NT) k_batch_from_n0f8(BatchArgs A) {
  __shared__ uint8_t tile[BA_CC][BA_T][BA_T + 1];
  const int tid = threadIdx.x, tiles = A.tiles_i * A.tiles_j;
  const int b = blockIdx.x / tiles, tl = blockIdx.x % tiles;
  const int i0 = (tl % A.tiles_i) * BA_T, j0 = (tl / A.tiles_i) * BA_T;
  const uint8_t* sp = A.src + (A.offsets ? A.offsets[b] : (long long)b * A.sb);
  const bool i_fast = A.si <= A.sj;
  for (int c0 = 0; c0 < A.C; c0 += BA_CC) {
    const int nc = min(BA_CC, A.C - c0);
    for (int e = tid; e < BA_T * BA_T * nc; e += BA_NT) {
      const int c = e % nc, r = e / nc;
      const int li = i_fast ? r % BA_T : r / BA_T, lj = i_fast ? r / BA_T : r % BA_T;
      const int gi = i0 + li, gj = j0 + lj;
      if (gi < A.M && gj < A.N) tile[c][lj][li] = sp[(c0 + c) * A.sc + gi * A.si + gj * A.sj];
    }
    __syncthreads();
    for (int e = tid; e < BA_T * BA_T * nc; e += BA_NT) {
      const int li = e % BA_T, lj = (e / BA_T) % BA_T, c = e / (BA_T * BA_T);
      const int gi = i0 + li, gj = j0 + lj;
      if (gi < A.M && gj < A.N)
        A.dst[gi + (size_t)A.M * (gj + (size_t)A.N * ((c0 + c) + (size_t)A.C * b))] = (float)tile[c][lj][li] / 255.f;
    }
    __syncthreads();
  }
}


// Fast paths of the same conversion for the two CONTIGUOUS layouts (every access a full, aligned vector; the general kernel
// above spends its time on per-element index arithmetic: 234 us for 64 x 512^2 x 3 samples against 40 us of DRAM time).
// (a) planar bytes in the destination's own order (stride_i = 1, stride_j = M, stride_c = M N, stride_b = M N C): a flat
//     elementwise pass, 16 samples per thread.
__global__ void __launch_bounds__(256) k_batch_flat_n0f8(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n16, size_t n) {
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n16; e += (size_t)gridDim.x * 256) {
    const uint4 v = reinterpret_cast<const uint4*>(src)[e];
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(dst) + 4 * e;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      o[q] = make_float4((float)(w[q] & 255u) / 255.f, (float)((w[q] >> 8) & 255u) / 255.f, (float)((w[q] >> 16) & 255u) / 255.f,
                         (float)(w[q] >> 24) / 255.f);
  }
  if (blockIdx.x == 0)   // tail (n not a multiple of 16)
    for (size_t e = 16 * n16 + threadIdx.x; e < n; e += 256) dst[e] = (float)src[e] / 255.f;
}
// (b) channel-interleaved pixels, dim 1 fastest (stride_c = 1, stride_i = C, stride_j = C M: Julia Matrix{RGB{N0f8}}): pixel
//     p = i + M j of image b sits at byte C p + c.  A thread takes 4 pixels = C aligned 32-bit words and writes one float4 per plane.
template <int C>
__global__ void __launch_bounds__(256) k_batch_interleaved_n0f8(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t px4,
                                                                size_t plane, long long sb, int B) {
  const size_t total = px4 * (size_t)B;   // px4 = M N / 4 groups of 4 pixels per image
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const size_t b = e / px4, g = e % px4;
    const unsigned* sp = reinterpret_cast<const unsigned*>(src + b * (size_t)sb) + g * C;
    unsigned w[C];
#pragma unroll
    for (int q = 0; q < C; ++q) w[q] = sp[q];
    float out[C][4];
#pragma unroll
    for (int k = 0; k < 4 * C; ++k)   // byte k of the group = channel k % C of pixel k / C
      out[k % C][k / C] = (float)((w[k / 4] >> (8 * (k % 4))) & 255u) / 255.f;
    float* dp = dst + b * (size_t)C * plane + 4 * g;
#pragma unroll
    for (int c = 0; c < C; ++c) *reinterpret_cast<float4*>(dp + (size_t)c * plane) = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
  }
}

}  // namespace admmtv
