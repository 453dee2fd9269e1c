// kernels_bwd.cuh -- hand-written adjoint of the unrolled iterations (replaces Zygote's tape
// through /root/reference/src/ops/ops.jl:166-174; recursion of SURVEY.md 8a-10 restated on the
// single state v_k = D x_k + u_{k-1}).
//
// With e(v) = v - ST(v) (= u), g(v) = ST(v) - e(v) (= z - u) the forward recursion is
//     x_k = A (b + rho D^T g(v_{k-1})),   v_k = D x_k + e(v_{k-1}),   A = F^-1 C F  (self-adjoint)
// and one backward iteration k = K..2, given vbar_k (cotangent of v_k; vbar_K = 0), is
//     xbar_k = [k==K] xbar + D^T vbar_k            (produced by the previous launch, in spectrum form)
//     rbar_k = A xbar_k                            k_dim2 (also G += Re(conj(F xbar_k) F r_k))
//     bbar  += rbar_k ;  gbar = rho D rbar_k ;  rhobar += <D rbar_k, g(v_{k-1})>
//     q = 2 gbar - vbar_k ;  m = 1[|v_{k-1}| > tau]
//     vbar_{k-1} = vbar_k - gbar + m q ;  taubar -= sum sign(v_{k-1}) m q
//     xbar_{k-1} = D^T vbar_{k-1}  -> dim-1 FFT                                   k_dim1_bwd
#pragma once

#include "kernels.cuh"

#ifndef ADMMTV_BWD_HOIST
#define ADMMTV_BWD_HOIST 1
#endif
#ifndef ADMMTV_MINB9B
#define ADMMTV_MINB9B 2
#endif
#ifndef ADMMTV_MINB8B
#define ADMMTV_MINB8B 2   // 256-row planes: 148 -> 128 registers, 2 blocks/SM: backward sweep 612 -> 443 us on 768 planes of 256^2
#endif
#ifndef ADMMTV_MINB7B
#define ADMMTV_MINB7B 4   // 128-row planes: 150 -> 136 us on 1024 planes of 128^2
#endif

namespace admmtv {

ADMMTV_DI float sgn_mask(float v, float tau) { return v > tau ? 1.f : (v < -tau ? -1.f : 0.f); }  // sign(v) * 1[|v|>tau]

struct BwdPoint {
  float2 vbar;  // vbar_{k-1}
};
// one (pixel, channel) of the adjoint update; d = (D rbar) component, v = v_{k-1}, eb = vbar_k.
// Accumulates the scalar partial sums when `own`.
ADMMTV_DI float2 bwd_point(float2 d, float2 v, float2 eb, float rho, float tau, bool own, double& racc, double& tacc) {
  const float2 sm = make_float2(sgn_mask(v.x, tau), sgn_mask(v.y, tau));
  const float2 gb = make_float2(rho * d.x, rho * d.y);
  const float2 q = make_float2(2.f * gb.x - eb.x, 2.f * gb.y - eb.y);
  const float2 m = make_float2(fabsf(sm.x), fabsf(sm.y));
  if (own) {
    const Shrunk s = shrink_aniso(v, tau);  // s.w = g(v)
    racc += (double)(d.x * s.w.x) + (double)(d.y * s.w.y);
    tacc -= (double)(sm.x * q.x) + (double)(sm.y * q.y);
  }
  return make_float2(eb.x - gb.x + m.x * q.x, eb.y - gb.y + m.y * q.y);
}

// isotropic: per-pixel shrink scale s, the coefficient of v in vbar (1[n>tau] tau ip / n^3) and the
// pixel's taubar term (1[n>tau] ip / n)
ADMMTV_DI void iso_pix(float nsq, float ip, float tau, float& s, float& coef, float& tterm) {
  const float n = sqrtf(nsq);
  s = n > 0.f ? fmaxf(1.f - tau / n, 0.f) : 0.f;
  const bool act = n > tau;
  coef = act ? tau * ip / (n * n * n) : 0.f;
  tterm = act ? ip / n : 0.f;
}
// one (pixel, channel) of the isotropic adjoint; returns vbar_{k-1}
ADMMTV_DI float2 iso_bwd_point(float2 d, float2 v, float2 eb, float rho, float s, float coef) {
  const float2 gb = make_float2(rho * d.x, rho * d.y);
  const float2 qq = make_float2(2.f * gb.x - eb.x, 2.f * gb.y - eb.y);
  return make_float2(eb.x - gb.x + s * qq.x + coef * v.x, eb.y - gb.y + s * qq.y + coef * v.y);
}
ADMMTV_DI float2 iso_bwd_full(float2 d, float2 v, float2 eb, float rho, float2 sc /* (s, coef) */, bool own, double& racc) {
  if (own) {
    const Shrunk g = shrink_iso(v, sc.x);
    racc += (double)(d.x * g.w.x) + (double)(d.y * g.w.y);
  }
  return iso_bwd_point(d, v, eb, rho, sc.x, sc.y);
}

// MODE 2: isotropic pass A -- dim-1 IFFT -> rbar_k ; bbar += rbar_k ; q = 2 rho D rbar_k - vbar_k ;
//         A.ip_out[q][pixel] = this pair's share of <q, v_{k-1}> (plain stores; k_iso_coef adds the pairs of a group in
//         a fixed order, so the result is bit-reproducible) ; nothing else is written.
// MODE 0: anisotropic.  MODE 1: isotropic pass B (per-pixel (s, tau ip / n^3) in A.sc from k_iso_coef, which also
// adds the per-pixel taubar terms; bbar was accumulated by pass A).
template <int LM, int MODE>
constexpr int kDim1BwdMinB = (LM == 9 && MODE == 0) ? ADMMTV_MINB9B : (LM == 8 ? ADMMTV_MINB8B : (LM == 7 ? ADMMTV_MINB7B : 1));

template <int LM, bool HAS_VBAR, int MODE, bool TMA>
ADMMTV_DI void dim1_bwd_body(const Dim1BwdArgs& A, const Dim1Tma* tm) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::COB, RPT = Cfg::RPT, CHUNK = Cfg::CHUNKB;
  static_assert(CO % CHUNK == 0, "chunking must divide the tile");
  ADMMTV_DYN_SMEM(float2, X);
  // 1-D grid: block = (pair q, column tile), tiles fastest (no 65535 limit on the number of pairs)
  const int tid = threadIdx.x, N = A.N, ntile = (A.N + CO - 1) / CO;
  const int bid = dim1_bid();
  const int q = bid / ntile;
  const int j0 = (bid % ntile) * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  const float2* sin_q = A.spec_in + (size_t)q * plane;
  auto jcol = [&](int c) {
    int j = j0 - 1 + c;
    if (j < 0) j += N;
    if (j >= N) j -= N;
    return j;
  };
  {
    const unsigned colb = (unsigned)(M * sizeof(float2));
    if (tid == 0) {
      l2_prefetch_bulk(A.vck + ((size_t)q * 2 + 0) * plane + (size_t)j0 * M, (unsigned)nout * colb);
      l2_prefetch_bulk(A.vck + ((size_t)q * 2 + 0) * plane + (size_t)jcol(nout + 1) * M, colb);
    }
    if (tid == 32 % NT) l2_prefetch_bulk(A.vck + ((size_t)q * 2 + 1) * plane + (size_t)j0 * M, (unsigned)nout * colb);
    if (HAS_VBAR && tid == 64 % NT) {
      l2_prefetch_bulk(A.vbar_in + ((size_t)q * 2 + 0) * plane + (size_t)j0 * M, (unsigned)nout * colb);
      l2_prefetch_bulk(A.vbar_in + ((size_t)q * 2 + 0) * plane + (size_t)jcol(nout + 1) * M, colb);
    }
    if (HAS_VBAR && tid == 96 % NT) l2_prefetch_bulk(A.vbar_in + ((size_t)q * 2 + 1) * plane + (size_t)j0 * M, (unsigned)nout * colb);
    if (!A.first && tid == 128 % NT) l2_prefetch_bulk(A.bbar + (size_t)q * plane + (size_t)j0 * M, (unsigned)nout * colb);
  }
  // 1. rbar_k for columns j0-1 .. j0+nout
#ifndef ADMMTV_EMU
  __shared__ unsigned long long tbar[1];
  if constexpr (TMA) {
    dim1_ifft_to_smem_tma<LM, NT>(X, MODE == 2 ? nout + 1 : nout + 2, &tm->in, [&](int c) { return q * N + jcol(c); }, tbar, A.twM, tid);
  } else
#endif
    dim1_ifft_to_smem<LM, NT>(X, MODE == 2 ? nout + 1 : nout + 2, [&](int c) { return sin_q + (size_t)jcol(c) * M; }, A.twM, tid);

  const int grp = q / A.pm.Qg;
  const float rho = A.rho[grp * A.PS + A.ir];                                     // rho_k
  const float tau = A.lambda[grp * A.PS + A.it] / A.rho[grp * A.PS + A.it];       // tau_{k-1}
  const int i0 = tid * RPT;
  double racc = 0.0, tacc = 0.0;
  const float2* sc_g = MODE == 1 ? A.sc + (size_t)grp * plane : nullptr;   // per-pixel (s, tau ip / n^3) from k_iso_coef
  auto PIX = [&](const float2* scp, size_t off, bool) { return scp[off]; };
  float* ip_o = MODE == 2 ? A.ip_out + (size_t)q * plane : nullptr;   // per-pair partial sums
  const float2* v1 = A.vck + ((size_t)q * 2 + 0) * plane;
  const float2* v2 = A.vck + ((size_t)q * 2 + 1) * plane;
  const float2* e1 = A.vbar_in + ((size_t)q * 2 + 0) * plane;
  const float2* e2 = A.vbar_in + ((size_t)q * 2 + 1) * plane;
  float2* o1 = A.vbar_out + ((size_t)q * 2 + 0) * plane;
  float2* o2 = A.vbar_out + ((size_t)q * 2 + 1) * plane;
  float2* bq = A.bbar + (size_t)q * plane;
  const float2 zero2 = make_float2(0.f, 0.f);

  float2 n1c[RPT];  // vbar_{k-1}, channel 1, current column
  float p1c[RPT];   // MODE 2: <q1, v1> of the current column
  if (MODE == 2) {
    const int j = jcol(1);
    float2 vv[RPT], ee[RPT];
    load_rows<RPT>(v1 + (size_t)j * M + i0, vv);
    if (HAS_VBAR) load_rows<RPT>(e1 + (size_t)j * M + i0, ee);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float2 d = csub(X[sidx<LM>(1, i0 + r)], X[sidx<LM>(0, i0 + r)]);
      float2 qq = make_float2(2.f * rho * d.x, 2.f * rho * d.y);
      if (HAS_VBAR) qq = csub(qq, ee[r]);
      p1c[r] = qq.x * vv[r].x + qq.y * vv[r].y;
    }
  } else {
    const int j = jcol(1);
    float2 vv[RPT], ee[RPT];
    float2 ss[RPT];
    load_rows<RPT>(v1 + (size_t)j * M + i0, vv);
    if (HAS_VBAR) load_rows<RPT>(e1 + (size_t)j * M + i0, ee);
    if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) ss[r] = PIX(sc_g, (size_t)j * M + i0 + r, false);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float2 d = csub(X[sidx<LM>(1, i0 + r)], X[sidx<LM>(0, i0 + r)]);
      if (MODE == 1) n1c[r] = iso_bwd_full(d, vv[r], HAS_VBAR ? ee[r] : zero2, rho, ss[r], true, racc);
      else n1c[r] = bwd_point(d, vv[r], HAS_VBAR ? ee[r] : zero2, rho, tau, true, racc, tacc);
    }
    store_rows<RPT>(o1 + (size_t)j * M + i0, n1c);
  }

  for (int c = 1; c <= nout; c += CHUNK) {
    // (a) anisotropic: every global load of the chunk is issued before any of them is consumed
    constexpr bool HOIST = ADMMTV_BWD_HOIST && MODE == 0 && RPT * CHUNK <= 4;  // larger tiles would spill
    float2 hv1[HOIST ? CHUNK : 1][RPT], he1[HOIST ? CHUNK : 1][RPT], hv2[HOIST ? CHUNK : 1][RPT + 1],
        he2[HOIST ? CHUNK : 1][RPT + 1], hbb[HOIST ? CHUNK : 1][RPT];
    if (HOIST) {
#pragma unroll
      for (int cc = 0; cc < CHUNK; ++cc) {
        const int j = jcol(c + cc), jn = jcol(c + cc + 1);
        load_rows<RPT>(v1 + (size_t)jn * M + i0, hv1[cc]);
        load_rows<RPT>(v2 + (size_t)j * M + i0, hv2[cc]);
        hv2[cc][RPT] = v2[(size_t)j * M + wrapm<M>(i0 + RPT)];
        if (HAS_VBAR) {
          load_rows<RPT>(e1 + (size_t)jn * M + i0, he1[cc]);
          load_rows<RPT>(e2 + (size_t)j * M + i0, he2[cc]);
          he2[cc][RPT] = e2[(size_t)j * M + wrapm<M>(i0 + RPT)];
        }
        if (!A.first) load_rows<RPT>(bq + (size_t)j * M + i0, hbb[cc]);
      }
    }
    float2 rr[CHUNK][RPT];
#pragma unroll
    for (int cc = 0; cc < CHUNK; ++cc) {
      const int col = c + cc;
      const int j = jcol(col), jn = jcol(col + 1);
      float2 xc[RPT + 2];  // rbar rows i0-1 .. i0+RPT of column col
      xc[0] = X[sidx<LM>(col, wrapm<M>(i0 - 1))];
#pragma unroll
      for (int r = 0; r < RPT; ++r) xc[r + 1] = X[sidx<LM>(col, i0 + r)];
      xc[RPT + 1] = X[sidx<LM>(col, wrapm<M>(i0 + RPT))];

      if (MODE == 2) {
        float2 bb[RPT], vv[RPT], ee[RPT], v2v[RPT], e2v[RPT];
        float p1n[RPT], ipv[RPT];
        if (!A.first) load_rows<RPT>(bq + (size_t)j * M + i0, bb);
        if (col < nout) {
          load_rows<RPT>(v1 + (size_t)jn * M + i0, vv);
          if (HAS_VBAR) load_rows<RPT>(e1 + (size_t)jn * M + i0, ee);
        }
        load_rows<RPT>(v2 + (size_t)j * M + i0, v2v);
        if (HAS_VBAR) load_rows<RPT>(e2 + (size_t)j * M + i0, e2v);
#pragma unroll
        for (int r = 0; r < RPT; ++r) bb[r] = A.first ? xc[r + 1] : cadd(bb[r], xc[r + 1]);
        store_rows<RPT>(bq + (size_t)j * M + i0, bb);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          if (col < nout) {
            const float2 d = csub(X[sidx<LM>(col + 1, i0 + r)], xc[r + 1]);
            float2 qq = make_float2(2.f * rho * d.x, 2.f * rho * d.y);
            if (HAS_VBAR) qq = csub(qq, ee[r]);
            p1n[r] = qq.x * vv[r].x + qq.y * vv[r].y;
          }
          const float2 d2 = csub(xc[r + 1], xc[r]);
          float2 q2 = make_float2(2.f * rho * d2.x, 2.f * rho * d2.y);
          if (HAS_VBAR) q2 = csub(q2, e2v[r]);
          ipv[r] = p1c[r] + q2.x * v2v[r].x + q2.y * v2v[r].y;
          p1c[r] = p1n[r];
        }
        store_rows_f<RPT>(ip_o + (size_t)j * M + i0, ipv);
        continue;
      }
      // bbar += rbar_k
      if (MODE == 0) {
        float2 bb[RPT];
        if (HOIST) {
#pragma unroll
          for (int r = 0; r < RPT; ++r) bb[r] = hbb[cc][r];
        } else if (!A.first) load_rows<RPT>(bq + (size_t)j * M + i0, bb);
#pragma unroll
        for (int r = 0; r < RPT; ++r) bb[r] = A.first ? xc[r + 1] : cadd(bb[r], xc[r + 1]);
        store_rows<RPT>(bq + (size_t)j * M + i0, bb);
      }
      // channel 1 at column col+1
      float2 n1n[RPT];
      {
        const bool own = col + 1 <= nout;
        float2 vv[RPT], ee[RPT];
        float2 ss[RPT];
        if (HOIST) {
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            vv[r] = hv1[cc][r];
            ee[r] = he1[cc][r];
          }
        } else {
          load_rows<RPT>(v1 + (size_t)jn * M + i0, vv);
          if (HAS_VBAR) load_rows<RPT>(e1 + (size_t)jn * M + i0, ee);
        }
        if (MODE == 1) {
#pragma unroll
          for (int r = 0; r < RPT; ++r) ss[r] = PIX(sc_g, (size_t)jn * M + i0 + r, false);
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const float2 d = csub(X[sidx<LM>(col + 1, i0 + r)], xc[r + 1]);
          if (MODE == 1) n1n[r] = iso_bwd_full(d, vv[r], HAS_VBAR ? ee[r] : zero2, rho, ss[r], own, racc);
          else n1n[r] = bwd_point(d, vv[r], HAS_VBAR ? ee[r] : zero2, rho, tau, own, racc, tacc);
        }
        if (own) store_rows<RPT>(o1 + (size_t)jn * M + i0, n1n);
      }
      // channel 2 at column col, rows i0 .. i0+RPT
      float2 n2[RPT + 1];
      {
        float2 vv[RPT + 1], ee[RPT + 1];
        float2 ss[RPT + 1];
        if (HOIST) {
#pragma unroll
          for (int r = 0; r <= RPT; ++r) {
            vv[r] = hv2[cc][r];
            ee[r] = he2[cc][r];
          }
        } else {
          load_rows<RPT>(v2 + (size_t)j * M + i0, vv);
          vv[RPT] = v2[(size_t)j * M + wrapm<M>(i0 + RPT)];
          if (HAS_VBAR) {
            load_rows<RPT>(e2 + (size_t)j * M + i0, ee);
            ee[RPT] = e2[(size_t)j * M + wrapm<M>(i0 + RPT)];
          }
        }
        if (MODE == 1) {
#pragma unroll
          for (int r = 0; r < RPT; ++r) ss[r] = PIX(sc_g, (size_t)j * M + i0 + r, true);   // own pixels: counted once
          ss[RPT] = PIX(sc_g, (size_t)j * M + wrapm<M>(i0 + RPT), false);
        }
#pragma unroll
        for (int r = 0; r <= RPT; ++r) {
          const float2 d = csub(xc[r + 1], xc[r]);
          if (MODE == 1) n2[r] = iso_bwd_full(d, vv[r], HAS_VBAR ? ee[r] : zero2, rho, ss[r], r < RPT, racc);
          else n2[r] = bwd_point(d, vv[r], HAS_VBAR ? ee[r] : zero2, rho, tau, r < RPT, racc, tacc);
        }
        store_rows<RPT>(o2 + (size_t)j * M + i0, n2);
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        rr[cc][r] = cadd(csub(n1c[r], n1n[r]), csub(n2[r], n2[r + 1]));  // xbar_{k-1} = D^T vbar_{k-1}
        n1c[r] = n1n[r];
      }
    }
    if (MODE == 2) continue;
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < CHUNK; ++cc)
#pragma unroll
      for (int r = 0; r < RPT; ++r) X[sidx<LM>(c + cc - 1, i0 + r)] = rr[cc][r];
  }
  if (MODE == 2) return;
  __syncthreads();

  float2* sout_q = A.spec_out + (size_t)q * plane;
#ifndef ADMMTV_EMU
  if constexpr (TMA) dim1_fft_from_smem_tma<LM, NT, false>(X, nout, &tm->out, q * N + j0, A.twM, tid);
  else
#endif
    dim1_fft_from_smem<LM, NT>(X, nout, [&](int c) { return sout_q + (size_t)(j0 + c) * M; }, A.twM, tid);

  // scalar partial sums -> fp64 accumulators (one atomic pair per block)
  const double rsum = block_sum(racc);
  const double tsum = block_sum(tacc);
  if (tid == 0) {
    atomicAdd(A.acc + (size_t)A.AS * grp + acc_rho(A.ir), rsum);
    atomicAdd(A.acc + (size_t)A.AS * grp + acc_tau(A.it), tsum);
#ifndef ADMMTV_EMU
    if constexpr (TMA) tma_store_wait_read();   // the TMA unit has read the tile: the block may exit
#endif
  }
}
template <int LM, bool HAS_VBAR, int MODE = 0>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT, kDim1BwdMinB<LM, MODE>) k_dim1_bwd(Dim1BwdArgs A) {
  dim1_bwd_body<LM, HAS_VBAR, MODE, false>(A, nullptr);
}
#ifndef ADMMTV_EMU
// the backward iteration kernel with TMA column loads / stores of the spectra (isotropic pass A only loads)
template <int LM, bool HAS_VBAR, int MODE = 0>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT, kDim1BwdMinB<LM, MODE>) k_dim1_bwd_tma(Dim1BwdArgs A, const __grid_constant__ Dim1Tma tm) {
  dim1_bwd_body<LM, HAS_VBAR, MODE, true>(A, &tm);
}
#endif

// Last backward iteration (k = 1): rbar_1 -> bbar_total = bbar + rbar_1, then either its dim-1
// FFT (MODE 0, feeds ybar = H bbar and the PSF correlation) or, with an empty PSF, ybar = bbar
// straight to the user layout (MODE 1).
template <int LM, int MODE, bool TMA>
ADMMTV_DI void dim1_bwd_last_body(const Dim1BwdArgs& A, const Dim1Tma* tm) {
  using Cfg = Dim1Cfg<LM>;
  constexpr int M = Cfg::M, NT = Cfg::NT, CO = Cfg::CO;
  ADMMTV_DYN_SMEM(float2, X);
  // 1-D grid: block = (pair q, column tile), tiles fastest (no 65535 limit on the number of pairs)
  const int tid = threadIdx.x, N = A.N, ntile = (A.N + CO - 1) / CO;
  const int q = blockIdx.x / ntile;
  const int j0 = (blockIdx.x % ntile) * CO;
  const int nout = min(CO, N - j0);
  const size_t plane = (size_t)N * M;
  const float2* sin_q = A.spec_in + (size_t)q * plane;
#ifndef ADMMTV_EMU
  __shared__ unsigned long long tbar[1];
  if constexpr (TMA) dim1_ifft_to_smem_tma<LM, NT>(X, nout, &tm->in, [&](int c) { return q * N + j0 + c; }, tbar, A.twM, tid);
  else
#endif
    dim1_ifft_to_smem<LM, NT>(X, nout, [&](int c) { return sin_q + (size_t)(j0 + c) * M; }, A.twM, tid);
  const float2* bq = A.bbar + (size_t)q * plane;
  const long ia = pm_in(A.pm, q, 0), ib = pm_in(A.pm, q, 1);  // ybar has y's layout
  const bool has_b = ib >= 0;
  for (int e = tid; e < nout * M; e += NT) {
    const int c = e / M, i = e % M;
    const size_t off = (size_t)(j0 + c) * M + i;
    float2 v = X[sidx<LM>(c, i)];
    if (!A.first) v = cadd(v, bq[off]);
    if (MODE == 0) X[sidx<LM>(c, i)] = v;
    else {
      if (A.pm.G > 1 && A.pm.in_gstride == 0) {   // groups share y: ybar (zeroed by the host) sums over groups
        atomicAdd(A.ybar + (size_t)ia * plane + off, v.x);
        if (has_b) atomicAdd(A.ybar + (size_t)ib * plane + off, v.y);
      } else {
        A.ybar[(size_t)ia * plane + off] = v.x;
        if (has_b) A.ybar[(size_t)ib * plane + off] = v.y;
      }
    }
  }
  if (MODE == 0) {
    __syncthreads();
    float2* sout_q = A.spec_out + (size_t)q * plane;
#ifndef ADMMTV_EMU
    if constexpr (TMA) dim1_fft_from_smem_tma<LM, NT>(X, nout, &tm->out, q * N + j0, A.twM, tid);
    else
#endif
      dim1_fft_from_smem<LM, NT>(X, nout, [&](int c) { return sout_q + (size_t)(j0 + c) * M; }, A.twM, tid);
  }
}
template <int LM, int MODE>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_dim1_bwd_last(Dim1BwdArgs A) {
  dim1_bwd_last_body<LM, MODE, false>(A, nullptr);
}
#ifndef ADMMTV_EMU
template <int LM, int MODE>
__global__ void __launch_bounds__(Dim1Cfg<LM>::NT) k_dim1_bwd_last_tma(Dim1BwdArgs A, const __grid_constant__ Dim1Tma tm) {
  dim1_bwd_last_body<LM, MODE, true>(A, &tm);
}
#endif

}  // namespace admmtv
