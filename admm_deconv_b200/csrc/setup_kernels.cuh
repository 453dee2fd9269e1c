// setup_kernels.cuh -- once-per-call table construction (included by admmtv_api.cu only).
#pragma once

#include "fft_core.cuh"
#include "args.cuh"
#include "launch_macros.cuh"
#include "reduce.cuh"

namespace admmtv {

// ------------------------------------------------------------------------------------------
// setup: twiddles, PSF spectrum, C / K tables (ops.jl:104-119 restated analytically)
// ------------------------------------------------------------------------------------------
// tw[n] = exp(-2 pi i n / L)
static __global__ void k_setup_twiddles(float2* twM, int M, float2* twN, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < M) {
    double s, c;
    sincospi(2.0 * n / M, &s, &c);
    twM[n] = make_float2((float)c, (float)(-s));
  }
  if (n < N) {
    double s, c;
    sincospi(2.0 * n / N, &s, &c);
    twN[n] = make_float2((float)c, (float)(-s));
  }
}

// clamp of deconv_admm.jl:216-219, in place; masks (1 = gradient passes) into `mask`:
// mask[0..PS) = lambda entries, mask[PS..2PS) = rho entries, mask[2PS..2PS+kh*kw) = h
static __global__ void k_clamp_params(float* lambda, float* rho, float* h, int nh, int PS, float creg, int do_clamp, float* mask) {
  // one block per group
  lambda += (size_t)blockIdx.x * PS;
  rho += (size_t)blockIdx.x * PS;
  if (h) h += (size_t)blockIdx.x * nh;
  if (mask) mask += (size_t)blockIdx.x * (nh + 2 * PS);
  for (int i = threadIdx.x; i < nh + 2 * PS; i += blockDim.x) {
    float* p = i < PS ? lambda + i : (i < 2 * PS ? rho + (i - PS) : h + (i - 2 * PS));
    const float v = *p;
    float m = 1.f;
    if (do_clamp) {
      if (i < 2 * PS) {
        m = (v >= creg) ? 1.f : 0.f;
        *p = (v != v) ? v : fmaxf(v, creg);   // Julia's clamp propagates NaN (fmaxf would replace it by the bound)
      } else {
        m = (v >= 0.f && v <= 1.f) ? 1.f : 0.f;
        *p = (v != v) ? v : fminf(fmaxf(v, 0.f), 1.f);
      }
    }
    if (mask) mask[i] = m;
  }
}

// T[k1][b] = sum_a h[a,b] exp(-2 pi i k1 a / M)      (dim-1 DFT of the corner-placed PSF)
static __global__ void k_setup_psf_dim1(const float* __restrict__ h, int kh, int kw, int M, double2* T) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * kw) return;
  h += (size_t)blockIdx.y * kh * kw;      // blockIdx.y = group
  T += (size_t)blockIdx.y * M * kw;
  const int k1 = idx / kw, b = idx % kw;
  double re = 0.0, im = 0.0;
  for (int a = 0; a < kh; ++a) {
    double s, c;
    sincospi(2.0 * ((long long)k1 * a % M) / M, &s, &c);
    const double hv = (double)h[a + kh * b];
    re += hv * c;
    im -= hv * s;
  }
  T[idx] = make_double2(re, im);
}

// ctab[p2][p1] = C(k1,k2) / (M N),  C = 1 / (|Sigma|^2 + rho (4 sin^2(pi k2/N) + 4 sin^2(pi k1/M)))   ops.jl:119
// ktab[p2][p1] = conj(K)(k1,k2) / N, K = Sigma exp(+2 pi i (k1 pd/M + k2 pr/N))                       SURVEY 8a-6
//   (only the dim-2 round trip's 1/N: the result of the k_dim2 pass is then the exact dim-1 spectrum of H^T y, which
//    feeds the first x-update directly; the dim-1 inverse that brings it back to space applies the remaining 1/M)
// sig [p2][p1] = Sigma (only written when sig != nullptr; the backward needs it)
// (p1,p2) are storage positions; (k1,k2) = pos_to_freq of them.
// blockIdx.z = parameter entry (per-iteration rho): one C table per entry, [G][PS][N][M]; K and Sigma do not depend on rho
static __global__ void k_setup_tables(const double2* __restrict__ T, int kh, int kw, int M, int N,
                                      const float* __restrict__ rho_p, float* ctab, float2* ktab, float2* sig, int planned) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  {  // blockIdx.y = group: its own rho, PSF spectrum scratch and tables
    const size_t grp = blockIdx.y, PS = gridDim.z, e = blockIdx.z;
    T += grp * (size_t)M * (kw > 0 ? kw : 1);
    rho_p += grp * PS + e;
    ctab += (grp * PS + e) * (size_t)M * N;
    if (e != 0) { ktab = nullptr; sig = nullptr; }
    if (ktab) ktab += grp * (size_t)M * N;
    if (sig) sig += grp * (size_t)M * N;
  }
  const int p1 = idx % M, p2 = idx / M;
  const int k1 = pos_to_freq(M, p1, (planned & 1) != 0), k2 = pos_to_freq(N, p2, (planned & 2) != 0, 1);
  double sr = 1.0, si = 0.0;
  if (kh > 0) {
    sr = 0.0;
    si = 0.0;
    for (int b = 0; b < kw; ++b) {
      double s, c;
      sincospi(2.0 * ((long long)k2 * b % N) / N, &s, &c);
      const double2 t = T[k1 * kw + b];
      // t * (c - i s)
      sr += t.x * c + t.y * s;
      si += t.y * c - t.x * s;
    }
  }
  const double rho = (double)*rho_p;
  double s1, c1, s2, c2;
  sincospi((double)k2 / N, &s2, &c2);
  sincospi((double)k1 / M, &s1, &c1);
  const double lap = 4.0 * s2 * s2 + 4.0 * s1 * s1;
  const double inv_mn = 1.0 / ((double)M * (double)N);
  const double C = 1.0 / (sr * sr + si * si + rho * lap);
  ctab[idx] = (float)(C * inv_mn);
  if (ktab) {
    const int pd = kh > 0 ? (kh - 1) / 2 : 0, pr = kw > 0 ? (kw - 1) / 2 : 0;
    double ps, pc;
    const double ph = 2.0 * ((double)((long long)k1 * pd % M) / M + (double)((long long)k2 * pr % N) / N);
    sincospi(ph, &ps, &pc);
    // K = Sigma * (pc + i ps); store conj(K)/(MN)
    const double kr = sr * pc - si * ps, ki = sr * ps + si * pc;
    const double inv_n = 1.0 / (double)N;
    ktab[idx] = make_float2((float)(kr * inv_n), (float)(-ki * inv_n));
  }
  if (sig) sig[idx] = make_float2((float)sr, (float)si);
}


// Isotropic TV (ops.jl:6,10): the per-pixel norm spans every plane of the call (a group).  The sweep kernels write each
// pair's share per pixel with plain stores; these kernels add the Qg shares of a group IN A FIXED ORDER (pair 0, 1, ...),
// so the norm -- and with it every threshold decision n > tau -- is bit-reproducible run to run (no float atomics).
//   part != null : nsq[g][i] = sum_q part[g*Qg + q][i]       (else nsq is taken as given: after a cross-rank all-reduce)
//   s_out != null: s_out[g][i] = max(1 - tau/n, 0), n = sqrt(nsq)
static __global__ void k_iso_scale(const float* __restrict__ part, int Qg, float* __restrict__ nsq, const float* __restrict__ lambda,
                                   const float* __restrict__ rho, int PS, int pe, float* __restrict__ s_out, int npix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (i >= npix) return;
  float t;
  if (part) {
    const float* pp = part + (size_t)g * Qg * npix + i;
    t = 0.f;
    for (int q = 0; q < Qg; ++q) t += pp[(size_t)q * npix];
    nsq[(size_t)g * npix + i] = t;
  } else {
    t = nsq[(size_t)g * npix + i];
  }
  if (s_out) {
    const float tau = lambda[g * PS + pe] / rho[g * PS + pe];   // parameter entry of this iteration
    const float n = sqrtf(t);
    s_out[(size_t)g * npix + i] = n > 0.f ? fmaxf(1.f - tau / n, 0.f) : 0.f;
  }
}
// backward: ip[g][i] = sum_q part[g*Qg + q][i] (fixed order; skipped when part is null), then, when sc != null,
// (s, 1[n>tau] tau ip / n^3) per pixel and taubar -= sum_pixels 1[n>tau] ip / n   (acc[8g+1], when count_tau)
static __global__ void k_iso_coef(const float* __restrict__ part, int Qg, const float* __restrict__ nsq, float* __restrict__ ip,
                                  const float* __restrict__ lambda, const float* __restrict__ rho, int PS, int pe, int AS,
                                  float2* __restrict__ sc, double* acc, int npix, int count_tau) {
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < npix;
  const int i = live ? i0 : 0;   // every thread takes part in the block reduction
  const int g = blockIdx.y;
  float p;
  if (part) {
    const float* pp = part + (size_t)g * Qg * npix + i;
    p = 0.f;
    for (int q = 0; q < Qg; ++q) p += pp[(size_t)q * npix];
    if (live) ip[(size_t)g * npix + i] = p;
  } else {
    p = ip[(size_t)g * npix + i];
  }
  if (!sc) return;   // uniform over the block
  const float tau = lambda[g * PS + pe] / rho[g * PS + pe];
  const float n = sqrtf(nsq[(size_t)g * npix + i]);
  const bool act = live && n > tau;
  if (live) {
    const float s = n > 0.f ? fmaxf(1.f - tau / n, 0.f) : 0.f;
    sc[(size_t)g * npix + i] = make_float2(s, act ? tau * p / (n * n * n) : 0.f);
  }
  const double tot = block_sum(act ? (double)(p / n) : 0.0);
  if (threadIdx.x == 0 && count_tau) atomicAdd(acc + (size_t)AS * g + acc_tau(pe), -tot);
}

// ------------------------------------------------------------------------------------------
// parameter gradients from the spectral accumulators ("after the loop")
//   Cbar = G/(MN) ; Sbar = -Cbar C^2 (cotangent of |Sigma|^2) ; rhobar += sum Sbar (|Lx|^2+|Ly|^2)
//   hbar[a,b] = Re sum_k W[k] e^{+2 pi i (k1 a/M + k2 b/N)},
//   W = 2 Sbar Sigma  +  (1/MN) P e^{-2 pi i (k1 pd/M + k2 pr/N)}   (spectral C path + spatial H^T y path)
// ------------------------------------------------------------------------------------------
// gacc and ctab hold one table per parameter entry ([G][PS][N][M]): Sbar_i from (G_i, C_i); rhobar_i gets its own spectral
// term; the cotangent of |Sigma|^2 is their sum.
static __global__ void k_grad_tables(const double* __restrict__ gacc, const float2* __restrict__ pacc,
                                     const float* __restrict__ ctab, const float2* __restrict__ sig, int kh, int kw,
                                     int M, int N, int use_spatial, double2* Wn, double* acc, int planned, int PS, int AS) {
  const int idx0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx0 < M * N;
  const int idx = live ? idx0 : 0;   // every thread takes part in the block reduction
  {  // blockIdx.y = group
    const size_t go = (size_t)blockIdx.y * M * N;
    gacc += go * PS; ctab += go * PS;
    if (pacc) pacc += go;
    if (sig) sig += go;
    if (Wn) Wn += go;
    acc += (size_t)AS * blockIdx.y;
  }
  const int p1 = idx % M, p2 = idx / M;
  const int k1 = pos_to_freq(M, p1, (planned & 1) != 0), k2 = pos_to_freq(N, p2, (planned & 2) != 0, 1);
  const double mn = (double)M * (double)N;
  double s1, c1, s2, c2;
  sincospi((double)k2 / N, &s2, &c2);
  sincospi((double)k1 / M, &s1, &c1);
  const double lap = 4.0 * s2 * s2 + 4.0 * s1 * s1;
  double Sbar = 0.0;
  for (int e = 0; e < PS; ++e) {
    const double C = (double)ctab[(size_t)e * M * N + idx] * mn;
    const double Se = -(gacc[(size_t)e * M * N + idx] / mn) * C * C;
    Sbar += Se;
    const double tot = block_sum(live ? Se * lap : 0.0);
    if (threadIdx.x == 0) atomicAdd(acc + acc_rhos(e), tot);
  }
  if (kh > 0 && live) {
    const float2 sg = sig[idx];
    double wr = 2.0 * Sbar * (double)sg.x, wi = 2.0 * Sbar * (double)sg.y;
    if (use_spatial) {
      const int pd = (kh - 1) / 2, pr = (kw - 1) / 2;
      double ps, pc;
      sincospi(2.0 * ((double)((long long)k1 * pd % M) / M + (double)((long long)k2 * pr % N) / N), &ps, &pc);
      const double pr_ = (double)pacc[idx].x / mn, pi_ = (double)pacc[idx].y / mn;
      // P * (pc - i ps)
      wr += pr_ * pc + pi_ * ps;
      wi += pi_ * pc - pr_ * ps;
    }
    Wn[(size_t)k2 * M + k1] = make_double2(wr, wi);
  }
}

// U[a][k2] = sum_k1 Wn[k2][k1] e^{+2 pi i k1 a / M}        one block per (a, k2): the M terms are
// spread over the block's threads (coalesced Wn row) and reduced in fp64
static __global__ void k_grad_h_dim1(const double2* __restrict__ Wn, int kh, int M, int N, double2* U) {
  const int a = blockIdx.x / N, k2 = blockIdx.x % N;
  Wn += (size_t)blockIdx.y * M * N;    // blockIdx.y = group
  U += (size_t)blockIdx.y * kh * N;
  double re = 0.0, im = 0.0;
  for (int k1 = threadIdx.x; k1 < M; k1 += blockDim.x) {
    double s, c;
    sincospi(2.0 * ((long long)k1 * a % M) / M, &s, &c);
    const double2 w = Wn[(size_t)k2 * M + k1];
    re += w.x * c - w.y * s;
    im += w.x * s + w.y * c;
  }
  re = block_sum(re);
  im = block_sum(im);
  if (threadIdx.x == 0) U[blockIdx.x] = make_double2(re, im);
  (void)kh;
}

// hbar[a,b] = mask * Re sum_k2 U[a][k2] e^{+2 pi i k2 b / N} (one block per tap); block 0 also
// finalises the scalars:
//   lambar = taubar / rho ; rhobar = direct + spectral - taubar lambda / rho^2   (tau = lambda ./ rho, ops.jl:102)
// acc: args.cuh acc_*.  mask: [0..PS) lambda, [PS..2PS) rho, [2PS..) h.
static __global__ void k_grad_finalize(const double2* __restrict__ U, int kh, int kw, int N, const float* __restrict__ mask,
                                       const double* __restrict__ acc, const float* __restrict__ lambda,
                                       const float* __restrict__ rho, float* hbar, float* lambar, float* rhobar,
                                       float* biasbar, int PS, int AS) {
  const int t = blockIdx.x;
  {  // blockIdx.y = group
    const int grp = blockIdx.y;
    U += (size_t)grp * kh * N;
    mask += (size_t)grp * (kh * kw + 2 * PS);
    acc += (size_t)AS * grp;
    lambda += (size_t)grp * PS; rho += (size_t)grp * PS; lambar += (size_t)grp * PS; rhobar += (size_t)grp * PS;
    if (hbar) hbar += (size_t)grp * kh * kw;
    if (biasbar) biasbar += grp;
  }
  if (t < kh * kw && hbar) {   // block-uniform
    const int a = t % kh, b = t / kh;
    double re = 0.0;
    for (int k2 = threadIdx.x; k2 < N; k2 += blockDim.x) {
      double s, c;
      sincospi(2.0 * ((long long)k2 * b % N) / N, &s, &c);
      const double2 u = U[a * N + k2];
      re += u.x * c - u.y * s;
    }
    re = block_sum(re);
    if (threadIdx.x == 0) hbar[t] = (float)(re * (double)mask[2 * PS + t]);
  }
  if (t == 0) {
    for (int e = threadIdx.x; e < PS; e += blockDim.x) {
      const double lam = (double)lambda[e], r = (double)rho[e], tb = acc[acc_tau(e)];
      lambar[e] = (float)((double)mask[e] * tb / r);
      rhobar[e] = (float)((double)mask[PS + e] * (acc[acc_rho(e)] + acc[acc_rhos(e)] - tb * lam / (r * r)));
    }
    if (biasbar && threadIdx.x == 0) *biasbar = (float)acc[acc_bias()];
  }
}

}  // namespace admmtv
