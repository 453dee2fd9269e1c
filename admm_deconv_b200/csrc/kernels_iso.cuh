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

#include "kernels.cuh"

namespace admmtv {

ADMMTV_DI float iso_scale(float nsq, float tau) {
  const float n = sqrtf(nsq);
  return n > 0.f ? fmaxf(1.f - tau / n, 0.f) : 0.f;  // ops.jl:10 (n = 0: max(-Inf,0)*0 = 0)
}
ADMMTV_DI Shrunk shrink_iso(float2 v, float s) {
  const float2 z = make_float2(s * v.x, s * v.y);
  Shrunk r;
  r.u = csub(v, z);
  r.w = csub(z, r.u);
  return r;
}

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
  const float rho = *A.rho;
  const float tau = *A.lambda / rho;
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
      const float s = iso_scale(A.nsq_in[off], tau);
      v1 = cadd(v1, shrink_iso(vp1[off], s).u);
      v2 = cadd(v2, shrink_iso(vp2[off], s).u);
    }
    vn1[off] = v1;
    vn2[off] = v2;
    atomicAdd(A.nsq_out + off, v1.x * v1.x + v1.y * v1.y + v2.x * v2.x + v2.y * v2.y);
  }
}

// ---- forward B ---------------------------------------------------------------------------------
template <int LM>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_iso_fwd_b(IsoArgs A) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::CO;
  ADMMTV_DYN_SMEM(float2, X);
  const int tid = threadIdx.x, N = A.N, q = blockIdx.y;
  const int j0 = blockIdx.x * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  const float rho = *A.rho;
  const float tau = *A.lambda / rho;
  const float2* v1 = A.v_in + ((size_t)q * 2 + 0) * plane;
  const float2* v2 = A.v_in + ((size_t)q * 2 + 1) * plane;
  const float2* bq = A.bpk + (size_t)q * plane;
  for (int e = tid; e < nout * M; e += NT) {
    const int c = e / M, i = e % M;
    const int j = j0 + c, jn = jwrap<LM>(j + 1, N), in = (i + 1) & (M - 1);
    const size_t o = (size_t)j * M + i, o_jn = (size_t)jn * M + i, o_in = (size_t)j * M + in;
    const float s = iso_scale(A.nsq_in[o], tau);
    const float2 w1 = shrink_iso(v1[o], s).w;
    const float2 w2 = shrink_iso(v2[o], s).w;
    const float2 w1n = shrink_iso(v1[o_jn], iso_scale(A.nsq_in[o_jn], tau)).w;
    const float2 w2n = shrink_iso(v2[o_in], iso_scale(A.nsq_in[o_in], tau)).w;
    const float2 dt = cadd(csub(w1, w1n), csub(w2, w2n));
    const float2 b = bq[o];
    X[sidx<LM>(c, i)] = make_float2(b.x + rho * dt.x, b.y + rho * dt.y);
  }
  __syncthreads();
  float2* sout_q = A.spec_out + (size_t)q * plane;
  dim1_fft_from_smem<LM, NT>(X, nout, [&](int c) { return sout_q + (size_t)(j0 + c) * M; }, A.twM, tid);
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
  const float rho = *A.rho;
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
    atomicAdd(A.ip + off, q1.x * a1.x + q1.y * a1.y + q2.x * a2.x + q2.y * a2.y);
  }
}

// one (pixel, channel) of the isotropic adjoint; returns vbar_{k-1}
ADMMTV_DI float2 iso_bwd_point(float2 d, float2 v, float2 eb, float rho, float s, float coef) {
  const float2 gb = make_float2(rho * d.x, rho * d.y);
  const float2 qq = make_float2(2.f * gb.x - eb.x, 2.f * gb.y - eb.y);
  return make_float2(eb.x - gb.x + s * qq.x + coef * v.x, eb.y - gb.y + s * qq.y + coef * v.y);
}

// ---- backward B --------------------------------------------------------------------------------
template <int LM, bool HAS_VBAR>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_iso_bwd_b(IsoArgs A) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::CO;
  ADMMTV_DYN_SMEM(float2, X);
  const int tid = threadIdx.x, N = A.N, q = blockIdx.y;
  const int j0 = blockIdx.x * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  const float2* sin_q = A.spec_in + (size_t)q * plane;
  // columns c = 0..nout+1  <->  j = j0-1+c
  dim1_ifft_to_smem<LM, NT>(X, nout + 2, [&](int c) { return sin_q + (size_t)jwrap<LM>(j0 - 1 + c, N) * M; }, A.twM, tid);
  const float rho = *A.rho;
  const float tau = *A.lambda / rho;
  const float2* v1 = A.v_in + ((size_t)q * 2 + 0) * plane;
  const float2* v2 = A.v_in + ((size_t)q * 2 + 1) * plane;
  const float2* e1 = A.vbar_in + ((size_t)q * 2 + 0) * plane;
  const float2* e2 = A.vbar_in + ((size_t)q * 2 + 1) * plane;
  float2* o1 = A.vbar_out + ((size_t)q * 2 + 0) * plane;
  float2* o2 = A.vbar_out + ((size_t)q * 2 + 1) * plane;
  const float2 zero2 = make_float2(0.f, 0.f);
  double racc = 0.0, tacc = 0.0;
  // per-pixel shrink scale s and the coefficient of v in vbar: 1[n>tau] tau ip / n^3
  auto pix = [&](size_t off, float& s, float& coef, float& tterm) {
    const float nsq = A.nsq_in[off];
    const float n = sqrtf(nsq);
    s = n > 0.f ? fmaxf(1.f - tau / n, 0.f) : 0.f;
    const bool act = n > tau;
    const float ip = A.ip[off];
    coef = act ? tau * ip / (n * n * n) : 0.f;
    tterm = act ? ip / n : 0.f;
  };
  constexpr int PER = (CO * M + NT - 1) / NT;
  float2 xbar_keep[PER];
#pragma unroll
  for (int cnt = 0; cnt < PER; ++cnt) {
    const int e = tid + cnt * NT;
    if (e >= nout * M) continue;
    const int c = e / M + 1, i = e % M;
    const int j = j0 + c - 1, jn = jwrap<LM>(j + 1, N), in = (i + 1) & (M - 1);
    const size_t o = (size_t)j * M + i, o_jn = (size_t)jn * M + i, o_in = (size_t)j * M + in;
    const float2 r = X[sidx<LM>(c, i)];
    const float2 r_im = X[sidx<LM>(c, (i - 1) & (M - 1))];
    const float2 r_ip = X[sidx<LM>(c, in)];
    const float2 r_jm = X[sidx<LM>(c - 1, i)];
    const float2 r_jp = X[sidx<LM>(c + 1, i)];
    float s, coef, tt, sj, coefj, ttj, si, coefi, tti;
    pix(o, s, coef, tt);
    pix(o_jn, sj, coefj, ttj);
    pix(o_in, si, coefi, tti);
    const float2 d1 = csub(r, r_jm), d2 = csub(r, r_im);
    const float2 a1 = v1[o], a2 = v2[o];
    const float2 n1 = iso_bwd_point(d1, a1, HAS_VBAR ? e1[o] : zero2, rho, s, coef);
    const float2 n2 = iso_bwd_point(d2, a2, HAS_VBAR ? e2[o] : zero2, rho, s, coef);
    const float2 n1n = iso_bwd_point(csub(r_jp, r), v1[o_jn], HAS_VBAR ? e1[o_jn] : zero2, rho, sj, coefj);
    const float2 n2n = iso_bwd_point(csub(r_ip, r), v2[o_in], HAS_VBAR ? e2[o_in] : zero2, rho, si, coefi);
    o1[o] = n1;
    o2[o] = n2;
    const Shrunk g1 = shrink_iso(a1, s), g2 = shrink_iso(a2, s);
    racc += (double)(d1.x * g1.w.x) + (double)(d1.y * g1.w.y) + (double)(d2.x * g2.w.x) + (double)(d2.y * g2.w.y);
    if (q == 0) tacc -= (double)tt;  // one term per PIXEL (the norm is shared by every plane)
    xbar_keep[cnt] = cadd(csub(n1, n1n), csub(n2, n2n));
  }
  __syncthreads();
#pragma unroll
  for (int cnt = 0; cnt < PER; ++cnt) {
    const int e = tid + cnt * NT;
    if (e >= nout * M) continue;
    const int c = e / M, i = e % M;
    X[sidx<LM>(c, i)] = xbar_keep[cnt];
  }
  __syncthreads();
  float2* sout_q = A.spec_out + (size_t)q * plane;
  dim1_fft_from_smem<LM, NT>(X, nout, [&](int c) { return sout_q + (size_t)(j0 + c) * M; }, A.twM, tid);
  const double rsum = block_sum(racc);
  const double tsum = block_sum(tacc);
  if (tid == 0) {
    atomicAdd(A.acc + 0, rsum);
    atomicAdd(A.acc + 1, tsum);
  }
}

}  // namespace admmtv
