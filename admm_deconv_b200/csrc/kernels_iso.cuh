// kernels_iso.cuh -- isotropic TV (block thresholding BT, ops.jl:6,10).
//
// BT couples every plane of the call: n[i,j] = sqrt(sum over dims 3,4 of the (M,N,2B,P) tensor of
// v^2) is ONE norm per pixel for both difference directions, all channels and all images of the
// batch.  That is a grid-wide reduction between computing v_k and shrinking it,
// so the fused dim-1 kernel splits in two around a per-pixel atomic accumulation:
//   forward   A: dim-1 IFFT -> x_k ; v_k = D x_k + u_{k-1}(n_{k-1}) ; store v_k ; nsq_k += |v_k|^2
//             B: s_k = max(1 - tau/n_k, 0) ; w = (2 s_k - 1) v_k ; r = b + rho D^T w ; dim-1 FFT
//   backward  A: dim-1 IFFT -> rbar_k ; bbar += rbar ; q = 2 rho D rbar - vbar_k ; ip += <q, v_{k-1}>
//             B: same IFFT again ; vbar_{k-1} = vbar_k - gbar + s q + 1[n>tau] tau ip v / n^3 ;
//                xbar_{k-1} = D^T vbar_{k-1} -> dim-1 FFT ; rhobar, taubar partial sums
// The float atomics make n (hence the result) reproducible only to rounding, not bitwise.
#pragma once

#include "kernels_bwd.cuh"

namespace admmtv {

template <int LM>
ADMMTV_DI int jwrap(int j, int N) {
  if (j < 0) j += N;
  if (j >= N) j -= N;
  return j;
}

// ---- forward A ---------------------------------------------------------------------------------
template <int LM, bool HAS_VPREV>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_iso_fwd_a(IsoArgs A) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::CO;
  ADMMTV_DYN_SMEM(float2, X);
  const int tid = threadIdx.x, N = A.N, q = blockIdx.y;
  const int j0 = blockIdx.x * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  const float2* sin_q = A.spec_in + (size_t)q * plane;
  // columns c = 0..nout  <->  j = j0-1+c
  dim1_ifft_to_smem<LM, NT>(X, nout + 1, [&](int c) { return sin_q + (size_t)jwrap<LM>(j0 - 1 + c, N) * M; }, A.twM, tid);
  const int grp = q / A.Qg;
  const float rho = A.rho[grp];
  const float tau = A.lambda[grp] / rho;
  const float* nsq_in = A.nsq_in + (size_t)grp * plane;
  float* nsq_out = A.nsq_out + (size_t)grp * plane;
  const float2* vp1 = A.v_in + ((size_t)q * 2 + 0) * plane;
  const float2* vp2 = A.v_in + ((size_t)q * 2 + 1) * plane;
  float2* vn1 = A.v_out + ((size_t)q * 2 + 0) * plane;
  float2* vn2 = A.v_out + ((size_t)q * 2 + 1) * plane;
  for (int e = tid; e < nout * M; e += NT) {
    const int c = e / M + 1, i = e % M;
    const size_t off = (size_t)(j0 + c - 1) * M + i;
    const float2 x = X[sidx<LM>(c, i)];
    float2 v1 = csub(x, X[sidx<LM>(c - 1, i)]);
    float2 v2 = csub(x, X[sidx<LM>(c, (i - 1) & (M - 1))]);
    if (HAS_VPREV) {
      const float s = iso_scale(nsq_in[off], tau);
      v1 = cadd(v1, shrink_iso(vp1[off], s).u);
      v2 = cadd(v2, shrink_iso(vp2[off], s).u);
    }
    vn1[off] = v1;
    vn2[off] = v2;
#ifdef ADMMTV_ISO_NOATOM   // timing experiment only
    if (v1.x == 12345.f) nsq_out[off] = 1.f;
#else
    atomicAdd(nsq_out + off, v1.x * v1.x + v1.y * v1.y + v2.x * v2.x + v2.y * v2.y);
#endif
  }
}

// ---- backward A --------------------------------------------------------------------------------
template <int LM, bool HAS_VBAR>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_iso_bwd_a(IsoArgs A) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::CO;
  ADMMTV_DYN_SMEM(float2, X);
  const int tid = threadIdx.x, N = A.N, q = blockIdx.y;
  const int j0 = blockIdx.x * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  const float2* sin_q = A.spec_in + (size_t)q * plane;
  dim1_ifft_to_smem<LM, NT>(X, nout + 1, [&](int c) { return sin_q + (size_t)jwrap<LM>(j0 - 1 + c, N) * M; }, A.twM, tid);
  const int grp = q / A.Qg;
  const float rho = A.rho[grp];
  float* ip_g = A.ip + (size_t)grp * plane;
  const float2* v1 = A.v_in + ((size_t)q * 2 + 0) * plane;     // v_{k-1}
  const float2* v2 = A.v_in + ((size_t)q * 2 + 1) * plane;
  const float2* e1 = A.vbar_in + ((size_t)q * 2 + 0) * plane;  // vbar_k
  const float2* e2 = A.vbar_in + ((size_t)q * 2 + 1) * plane;
  float2* bq = A.bbar + (size_t)q * plane;
  for (int e = tid; e < nout * M; e += NT) {
    const int c = e / M + 1, i = e % M;
    const size_t off = (size_t)(j0 + c - 1) * M + i;
    const float2 r = X[sidx<LM>(c, i)];
    bq[off] = A.first ? r : cadd(bq[off], r);
    const float2 d1 = csub(r, X[sidx<LM>(c - 1, i)]);
    const float2 d2 = csub(r, X[sidx<LM>(c, (i - 1) & (M - 1))]);
    float2 q1 = make_float2(2.f * rho * d1.x, 2.f * rho * d1.y);
    float2 q2 = make_float2(2.f * rho * d2.x, 2.f * rho * d2.y);
    if (HAS_VBAR) {
      q1 = csub(q1, e1[off]);
      q2 = csub(q2, e2[off]);
    }
    const float2 a1 = v1[off], a2 = v2[off];
    atomicAdd(ip_g + off, q1.x * a1.x + q1.y * a1.y + q2.x * a2.x + q2.y * a2.y);
  }
}

}  // namespace admmtv
