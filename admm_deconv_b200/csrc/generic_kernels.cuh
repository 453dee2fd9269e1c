// generic_kernels.cuh -- the ADMM-TV path for image sizes that have no register-FFT plan (fft_core.cuh).
//
// The reference accepts every (M,N) because FFTW does (/root/reference/src/ops/ops.jl:26,86).  The tuned kernels
// of kernels.cuh / kernels_bwd.cuh exist for the 20 planned lengths; any other size 1..4096 runs through the
// kernels below, which implement the SAME launcher contracts (args.cuh: Dim1Launch / Dim2Launch, size id 0)
// with run-time sizes, so the launch sequences, workspace layout, checkpoint and gradient code of
// admmtv_api.cu / bwd_api.inc are shared.  Differences from the tuned path:
//   * spectra are in NATURAL order (pos_to_freq is the identity for unplanned lengths);
//   * a length-L DFT is a two- or three-factor Cooley-Tukey decomposition L = L1*La*Lb (the split minimising L1+La+Lb)
//     with direct sub-transforms in shared memory: O(L (L1+La+Lb)) per line, O(L^2) for a prime length;
//   * the stencil sweeps are separate, unfused per-pixel kernels (neighbour values are recomputed, not exchanged).
// Same arithmetic per pixel (shrink_aniso / bwd_point / iso_* are the tuned path's functions), same results
// to fp32 rounding; correct and bandwidth-reasonable, not roofline-tuned.
#pragma once

#include "kernels_bwd.cuh"

namespace admmtv {

constexpr int GK_NT = 256;

ADMMTV_HD inline int small_factor(int L) {
  int f = 1;
  for (int d = 1; d * d <= L; ++d)
    if (L % d == 0) f = d;
  return f;
}
// L = L1 * La * Lb minimising L1 + La + Lb (the number of complex multiply-adds per element); La = 1 means a two-factor plan
// L = L1 * Lb.  L1 <= La <= Lb, so a prime gives (1, 1, L).
struct DftPlan {
  int L1, La, Lb;
};
ADMMTV_HD inline DftPlan dft_plan(int L) {
  DftPlan best{1, 1, L};
  int cost = L + 2;
  for (int a = 1; a * a * a <= L; ++a) {
    if (L % a) continue;
    const int r = L / a;
    for (int b = a; b * b <= r; ++b) {
      if (r % b) continue;
      const int c = r / b;
      const int cst = a + b + c - (a == 1 ? 1 : 0) - (b == 1 ? 1 : 0);   // a factor of 1 is no pass at all
      if (cst < cost) {
        cost = cst;
        best = DftPlan{a, b, c};
      }
    }
  }
  if (best.L1 == 1 && best.La > 1) {   // two factors: make L1 the small one, no inner split
    best.L1 = best.La;
    best.La = 1;
  }
  return best;
}
ADMMTV_HD inline int gk_lines_dim1(int M) { return M <= 64 ? 8 : (M <= 256 ? 4 : (M <= 1024 ? 2 : 1)); }
ADMMTV_HD inline int gk_rows_dim2(int N) { return N <= 1024 ? 8 : (N <= 2048 ? 4 : 2); }

ADMMTV_DI float2 cfma(float2 acc, float2 a, float2 w) {
  return make_float2(acc.x + a.x * w.x - a.y * w.y, acc.y + a.x * w.y + a.y * w.x);
}

// DFT of `nl` lines xs[l*L .. l*L+L) held in shared memory (natural order in and out) with scratch ys of the same size;
// RETURNS the buffer that holds the result (xs for a two-factor plan, ys for a three-factor one).  L = L1 * L2, L2 = La * Lb.
// tw[n] = exp(-2 pi i n / L), twb[m] = tw[L1 m] (m < L2), twc[m] = twb[La m] (m < Lb), all in shared memory; `inv` conjugates
// the twiddles (no 1/L: the tables carry the normalisation).  Every thread of the block must call; ends with a barrier.
// Thread mappings keep shared-memory reads contiguous or broadcast within a warp.
ADMMTV_DI float2* line_dft(float2* xs, float2* ys, const float2* tw, const float2* twb, const float2* twc, int L, DftPlan P, int nl,
                           bool inv, int tid) {
  const float sg = inv ? -1.f : 1.f;
  const int L1 = P.L1, L2 = P.La * P.Lb;
  // ys[k1*L2 + n2] = W_L^(n2 k1) * sum_n1 xs[L2 n1 + n2] W_L^(L2 n1 k1)
  for (int e = tid; e < nl * L; e += GK_NT) {
    const int l = e / L, r = e % L, k1 = r / L2, n2 = r % L2;
    const float2* x = xs + l * L + n2;
    float2 acc = make_float2(0.f, 0.f);
    const int step = (int)(((long long)L2 * k1) % L);
    int idx = 0;
    for (int n1 = 0; n1 < L1; ++n1) {
      const float2 w = tw[idx];
      acc = cfma(acc, x[n1 * L2], make_float2(w.x, sg * w.y));
      idx += step;
      if (idx >= L) idx -= L;
    }
    const float2 w = tw[(int)(((long long)n2 * k1) % L)];
    ys[e] = cmul(acc, make_float2(w.x, sg * w.y));
  }
  __syncthreads();
  if (P.La <= 1) {
    // xs[k1 + L1 k2] = sum_n2 ys[k1*L2 + n2] W_L2^(n2 k2),  W_L2^m = twb[m]
    for (int e = tid; e < nl * L; e += GK_NT) {
      const int l = e / L, r = e % L, k1 = r / L2, k2 = r % L2;
      const float2* y = ys + l * L + k1 * L2;
      float2 acc = make_float2(0.f, 0.f);
      int m = 0;
      for (int n2 = 0; n2 < L2; ++n2) {
        const float2 w = twb[m];
        acc = cfma(acc, y[n2], make_float2(w.x, sg * w.y));
        m += k2;
        if (m >= L2) m -= L2;
      }
      xs[l * L + k1 + L1 * k2] = acc;
    }
    __syncthreads();
    return xs;
  }
  // three factors: the nl*L1 length-L2 lines of ys get the same two-step treatment with (La, Lb), scratch = xs
  const int La = P.La, Lb = P.Lb;
  //   xs[lam*L2 + ka*Lb + nb] = W_L2^(nb ka) * sum_na ys[lam*L2 + Lb na + nb] W_L2^(Lb na ka)
  for (int e = tid; e < nl * L; e += GK_NT) {
    const int lam = e / L2, r = e % L2, ka = r / Lb, nb = r % Lb;
    const float2* y = ys + lam * L2 + nb;
    float2 acc = make_float2(0.f, 0.f);
    const int step = (Lb * ka) % L2;
    int idx = 0;
    for (int na = 0; na < La; ++na) {
      const float2 w = twb[idx];
      acc = cfma(acc, y[na * Lb], make_float2(w.x, sg * w.y));
      idx += step;
      if (idx >= L2) idx -= L2;
    }
    const float2 w = twb[(nb * ka) % L2];
    xs[e] = cmul(acc, make_float2(w.x, sg * w.y));
  }
  __syncthreads();
  //   result[k1 + L1 (ka + La kb)] = sum_nb xs[lam*L2 + ka*Lb + nb] W_Lb^(nb kb),  W_Lb^m = twc[m];  lam = l*L1 + k1
  for (int e = tid; e < nl * L; e += GK_NT) {
    const int lam = e / L2, r = e % L2, ka = r / Lb, kb = r % Lb;
    const float2* x = xs + lam * L2 + ka * Lb;
    float2 acc = make_float2(0.f, 0.f);
    int m = 0;
    for (int nb = 0; nb < Lb; ++nb) {
      const float2 w = twc[m];
      acc = cfma(acc, x[nb], make_float2(w.x, sg * w.y));
      m += kb;
      if (m >= Lb) m -= Lb;
    }
    const int l = lam / L1, k1 = lam % L1;
    ys[l * L + k1 + L1 * (ka + La * kb)] = acc;
  }
  __syncthreads();
  return ys;
}

// ---- dim-1 DFT of contiguous lines, in place or out of place ------------------------------------------
struct GDft1Args {
  const float2* in;
  float2* out;
  const float2* tw;
  int M, LB, inv;
  DftPlan P;
  long long nlines;
};
ADMMTV_HD inline size_t gk_dft_smem_elems(int L, DftPlan P, int lines) { return (size_t)L + P.La * P.Lb + P.Lb + 2 * (size_t)lines * L; }
__global__ void __launch_bounds__(GK_NT) gk_dft1(GDft1Args A) {
  ADMMTV_DYN_SMEM(float2, sm);
  const int L2 = A.P.La * A.P.Lb;
  float2* tw = sm;
  float2* twb = tw + A.M;
  float2* twc = twb + L2;
  float2* xs = twc + A.P.Lb;
  float2* ys = xs + (size_t)A.LB * A.M;
  const int tid = threadIdx.x;
  const long long l0 = (long long)blockIdx.x * A.LB;
  const int nl = (int)(A.nlines - l0 < A.LB ? A.nlines - l0 : A.LB);
  for (int e = tid; e < A.M; e += GK_NT) tw[e] = A.tw[e];
  for (int e = tid; e < L2; e += GK_NT) twb[e] = A.tw[e * A.P.L1];
  for (int e = tid; e < A.P.Lb; e += GK_NT) twc[e] = A.tw[e * A.P.L1 * A.P.La];
  for (int e = tid; e < nl * A.M; e += GK_NT) xs[e] = A.in[l0 * A.M + e];
  __syncthreads();
  const float2* res = line_dft(xs, ys, tw, twb, twc, A.M, A.P, nl, A.inv != 0, tid);
  for (int e = tid; e < nl * A.M; e += GK_NT) A.out[l0 * A.M + e] = res[e];
}

// ---- dim-2: forward DFT -> [save Z] -> [accumulate] -> x table -> inverse DFT  (the contract of k_dim2) -----
struct GDim2Cfg {
  int N, TR, mul, save_z, acc, fwd_only;
  DftPlan P;
};
__global__ void __launch_bounds__(GK_NT) gk_dim2(Dim2Args A, GDim2Cfg C) {
  ADMMTV_DYN_SMEM(float2, sm);
  const int N2 = C.P.La * C.P.Lb;
  float2* tw = sm;
  float2* twb = tw + C.N;
  float2* twc = twb + N2;
  float2* xs = twc + C.P.Lb;
  float2* ys = xs + (size_t)C.TR * C.N;
  const int tid = threadIdx.x, M = A.M, N = C.N, TR = C.TR;
  const int row_tiles = (M + TR - 1) / TR;
  const int q = blockIdx.x / row_tiles, i0 = (blockIdx.x % row_tiles) * TR;
  const int nr = min(TR, M - i0);
  const size_t qoff = (size_t)q * N * M;
  const size_t toff = (size_t)(q / A.Qg) * A.tab_stride;
  for (int e = tid; e < N; e += GK_NT) tw[e] = A.twN[e];
  for (int e = tid; e < N2; e += GK_NT) twb[e] = A.twN[e * C.P.L1];
  for (int e = tid; e < C.P.Lb; e += GK_NT) twc[e] = A.twN[e * C.P.L1 * C.P.La];
  for (int e = tid; e < nr * N; e += GK_NT) {
    const int ll = e % nr, col = e / nr;
    xs[ll * N + col] = A.in[qoff + (size_t)col * M + i0 + ll];
  }
  __syncthreads();
  float2* zs = line_dft(xs, ys, tw, twb, twc, N, C.P, nr, false, tid);   // spectrum lines (xs or ys)
  float2* sc = zs == xs ? ys : xs;                                        // the other buffer: scratch of the inverse
  for (int e = tid; e < nr * N; e += GK_NT) {
    const int ll = e % nr, col = e / nr;
    const size_t g = (size_t)col * M + i0 + ll;
    float2 z = zs[ll * N + col];
    if (C.save_z || C.fwd_only) (C.fwd_only ? A.out : A.zsave)[qoff + g] = z;
    if (C.acc) {
      const float2 z2 = A.z2[qoff + g];
      const float re = z.x * z2.x + z.y * z2.y, im = z.x * z2.y - z.y * z2.x;  // conj(Z) Z2
      if (C.acc == 1) atomicAdd(A.gacc + toff + g, (double)re);
      else {
        atomicAdd(reinterpret_cast<float*>(A.pacc + toff + g), re);
        atomicAdd(reinterpret_cast<float*>(A.pacc + toff + g) + 1, im);
      }
    }
    if (!C.fwd_only) {
      if (C.mul == 0) z = cscale(z, A.ctab[toff + g]);
      else {
        const float2 kk = A.ktab[toff + g];
        z = cmul(z, make_float2(kk.x, C.mul == 2 ? -kk.y : kk.y));
      }
      zs[ll * N + col] = z;
    }
  }
  if (C.fwd_only) return;
  __syncthreads();
  const float2* res = line_dft(zs, sc, tw, twb, twc, N, C.P, nr, true, tid);
  for (int e = tid; e < nr * N; e += GK_NT) {
    const int ll = e % nr, col = e / nr;
    A.out[qoff + (size_t)col * M + i0 + ll] = res[ll * N + col];
  }
}

// ---- per-pixel helpers ---------------------------------------------------------------------------------
struct Pix {
  int i, j, q, ip1, im1, jp1, jm1;  // circular neighbours
  size_t plane, o;                  // o = j*M + i
};
ADMMTV_DI bool gk_pix(size_t idx, int M, int N, int Q, Pix& p) {
  p.plane = (size_t)M * N;
  if (idx >= p.plane * Q) return false;
  p.i = (int)(idx % M);
  p.j = (int)((idx / M) % N);
  p.q = (int)(idx / p.plane);
  p.ip1 = p.i + 1 == M ? 0 : p.i + 1;
  p.im1 = p.i == 0 ? M - 1 : p.i - 1;
  p.jp1 = p.j + 1 == N ? 0 : p.j + 1;
  p.jm1 = p.j == 0 ? N - 1 : p.j - 1;
  p.o = (size_t)p.j * M + p.i;
  return true;
}

// ---- pack (the elementwise part of k_pack_fft1; the dim-1 DFT follows as gk_dft1 in place) -------------------
__global__ void __launch_bounds__(GK_NT) gk_pack(PackArgs A, int M, int Q, int mode) {
  const size_t idx = (size_t)blockIdx.x * GK_NT + threadIdx.x;
  Pix p;
  double bsum = 0.0, lsum = 0.0;
  int grp = 0;
  if (gk_pix(idx, M, A.N, Q, p)) {
    grp = p.q / A.pm.Qg;
    if (mode == 2) A.spec[idx] = A.src_packed[idx];
    else {
      const bool cot = mode == 1 || mode == 3;
      const long ia = cot ? pm_out(A.pm, p.q, 0) : pm_in(A.pm, p.q, 0);
      const long ib = cot ? pm_out(A.pm, p.q, 1) : pm_in(A.pm, p.q, 1);
      const float* srcp = mode == 3 ? A.target : A.src;
      float va = srcp[(size_t)ia * p.plane + p.o], vb = ib >= 0 ? srcp[(size_t)ib * p.plane + p.o] : 0.f;
      if (mode == 0 && A.packed_out) A.packed_out[idx] = make_float2(va, vb);
      if (mode == 3) {   // xbar = mse_scale (x_out - target), loss += (x_out - target)^2
        const float da = A.xout[(size_t)ia * p.plane + p.o] - va, db = ib >= 0 ? A.xout[(size_t)ib * p.plane + p.o] - vb : 0.f;
        lsum = (double)(da * da + db * db);
        va = A.mse_scale * da; vb = A.mse_scale * db;
      }
      if (cot) {
        va *= act_grad_from_out(A.xout[(size_t)ia * p.plane + p.o], A.act);
        if (ib >= 0) vb *= act_grad_from_out(A.xout[(size_t)ib * p.plane + p.o], A.act);
        bsum = (double)va + (double)vb;
      }
      A.spec[idx] = make_float2(va, vb);
    }
  }
  if (mode == 3) {
    const double tot = block_sum(lsum);
    if (threadIdx.x == 0) atomicAdd(A.loss_acc, tot);
  }
  if ((mode == 1 || mode == 3) && A.bias_acc) {
    // one atomic per block when the whole block lies in one group, else (at most G-1 blocks) one per thread
    const size_t per_group = (size_t)A.pm.Qg * M * A.N, first = (size_t)blockIdx.x * GK_NT;
    size_t last = first + GK_NT - 1;
    if (last >= per_group * A.pm.G) last = per_group * A.pm.G - 1;
    if (first / per_group == last / per_group) {
      const double tot = block_sum(bsum);
      if (threadIdx.x == 0) atomicAdd(A.bias_acc + (size_t)A.acc_stride * (int)(first / per_group), tot);
    } else if (bsum != 0.0) {
      atomicAdd(A.bias_acc + (size_t)A.acc_stride * grp, bsum);
    }
  }
}

// ---- out (the elementwise part of k_dim1_out; the inverse dim-1 DFT ran in place before) ---------------------
__global__ void __launch_bounds__(GK_NT) gk_out(OutArgs A, int M, int Q, int mode) {
  const size_t idx = (size_t)blockIdx.x * GK_NT + threadIdx.x;
  Pix p;
  if (!gk_pix(idx, M, A.N, Q, p)) return;
  float2 v = A.spec[idx];
  if (mode != 1) v = cscale(v, A.scale);
  if (mode == 0) {
    A.packed[idx] = v;
  } else if (mode == 2) {
    const long ia = pm_in(A.pm, p.q, 0), ib = pm_in(A.pm, p.q, 1);
    if (A.pm.G > 1 && A.pm.in_gstride == 0) {
      atomicAdd(A.planes + (size_t)ia * p.plane + p.o, v.x);
      if (ib >= 0) atomicAdd(A.planes + (size_t)ib * p.plane + p.o, v.y);
    } else {
      A.planes[(size_t)ia * p.plane + p.o] = v.x;
      if (ib >= 0) A.planes[(size_t)ib * p.plane + p.o] = v.y;
    }
  } else {
    const long ia = pm_out(A.pm, p.q, 0), ib = pm_out(A.pm, p.q, 1);
    const float bias = A.bias ? A.bias[p.q / A.pm.Qg] : 0.f;
    A.planes[(size_t)ia * p.plane + p.o] = act_apply(v.x + bias, A.act);
    if (ib >= 0) A.planes[(size_t)ib * p.plane + p.o] = act_apply(v.y + bias, A.act);
  }
}

// ---- forward sweep (the stencil of k_dim1_fwd; X = x_k, spatial, [Q][N][M]) --------------------------------------
//   mode 0: v_k = D x_k + u_{k-1} (stored) ; r = b + rho D^T (z_k - u_k) -> rout           (ops.jl:168-173)
//   mode 2: isotropic pass A: v_k stored, nsq_out[q][pixel] = this pair's |v_k|^2 (summed in order by k_iso_scale)
//   mode 1: isotropic pass B: r = b + rho D^T ((2 s_k - 1) v_k) -> rout   (v_k in A.vprev, s_k in A.nsq)
__global__ void __launch_bounds__(GK_NT) gk_sweep_fwd(Dim1FwdArgs A, const float2* X, float2* rout, int M, int Q, int mode,
                                                      int has_vprev) {
  const size_t idx = (size_t)blockIdx.x * GK_NT + threadIdx.x;
  Pix p;
  if (!gk_pix(idx, M, A.N, Q, p)) return;
  const int grp = p.q / A.Qg;
  const float rho = A.rho[grp * A.PS + A.in], tau = A.lambda[grp * A.PS + A.ic] / A.rho[grp * A.PS + A.ic];
  const float tau_p = A.lambda[grp * A.PS + A.ip] / A.rho[grp * A.PS + A.ip];
  const size_t q1 = ((size_t)p.q * 2 + 0) * p.plane, q2 = ((size_t)p.q * 2 + 1) * p.plane, qx = (size_t)p.q * p.plane;
  const size_t o = p.o, oj = (size_t)p.jp1 * M + p.i, oi = (size_t)p.j * M + p.ip1;   // (i,j), (i,j+1), (i+1,j)
  const float* ng = mode != 0 ? A.nsq + (size_t)grp * p.plane : nullptr;
  auto SC = [&](float t) { return t; };   // A.nsq holds the shrink scale s (k_iso_scale)
  if (mode == 1) {
    const float2 w1 = shrink_iso(A.vprev[q1 + o], SC(ng[o])).w, w1n = shrink_iso(A.vprev[q1 + oj], SC(ng[oj])).w;
    const float2 w2 = shrink_iso(A.vprev[q2 + o], SC(ng[o])).w, w2n = shrink_iso(A.vprev[q2 + oi], SC(ng[oi])).w;
    const float2 dt = cadd(csub(w1, w1n), csub(w2, w2n));
    const float2 b = A.bpk[qx + o];
    rout[qx + o] = make_float2(b.x + rho * dt.x, b.y + rho * dt.y);
    return;
  }
  const float2 x0 = X[qx + o];
  // u_{k-1} at a pixel / channel
  auto U = [&](size_t base, size_t off) {
    if (!has_vprev) return make_float2(0.f, 0.f);
    const float2 vp = A.vprev[base + off];
    return mode == 2 ? shrink_iso(vp, SC(ng[off])).u : shrink_aniso(vp, tau_p).u;
  };
  const float2 v1 = cadd(csub(x0, X[qx + (size_t)p.jm1 * M + p.i]), U(q1, o));   // channel 1: dim-2 difference
  const float2 v2 = cadd(csub(x0, X[qx + (size_t)p.j * M + p.im1]), U(q2, o));   // channel 2: dim-1 difference
  A.vnew[q1 + o] = v1;
  A.vnew[q2 + o] = v2;
  if (mode == 2) {
    A.nsq_out[qx + o] = v1.x * v1.x + v1.y * v1.y + v2.x * v2.x + v2.y * v2.y;
    return;
  }
  const float2 v1n = cadd(csub(X[qx + oj], x0), U(q1, oj));   // channel 1 at (i, j+1)
  const float2 v2n = cadd(csub(X[qx + oi], x0), U(q2, oi));   // channel 2 at (i+1, j)
  const float2 dt = cadd(csub(shrink_aniso(v1, tau).w, shrink_aniso(v1n, tau).w),
                         csub(shrink_aniso(v2, tau).w, shrink_aniso(v2n, tau).w));   // D^T (z - u)
  const float2 b = A.bpk[qx + o];
  rout[qx + o] = make_float2(b.x + rho * dt.x, b.y + rho * dt.y);
}

// ---- backward sweep (the stencil of k_dim1_bwd; R = rbar_k, spatial) ---------------------------------------------
//   mode 0 / 1: bbar (+)= R (mode 0) ; vbar_{k-1} stored ; xbar_{k-1} = D^T vbar_{k-1} -> xout ; rhobar / taubar sums
//   mode 2: isotropic pass A: bbar (+)= R ; ip_out[q][pixel] = this pair's <q, v_{k-1}> (summed in order by k_iso_coef)
__global__ void __launch_bounds__(GK_NT) gk_sweep_bwd(Dim1BwdArgs A, const float2* R, float2* xout, int M, int Q, int mode,
                                                      int has_vbar) {
  const size_t idx = (size_t)blockIdx.x * GK_NT + threadIdx.x;
  Pix p;
  const bool live = gk_pix(idx, M, A.N, Q, p);
  double racc = 0.0, tacc = 0.0;
  int grp = 0;
  if (live) {
    grp = p.q / A.pm.Qg;
    const float rho = A.rho[grp * A.PS + A.ir], tau = A.lambda[grp * A.PS + A.it] / A.rho[grp * A.PS + A.it];
    const size_t q1 = ((size_t)p.q * 2 + 0) * p.plane, q2 = ((size_t)p.q * 2 + 1) * p.plane, qx = (size_t)p.q * p.plane;
    const size_t o = p.o, oj = (size_t)p.jp1 * M + p.i, oi = (size_t)p.j * M + p.ip1;
    const float2 zero2 = make_float2(0.f, 0.f);
    const float2 r0 = R[qx + o];
    auto E = [&](size_t base, size_t off) { return has_vbar ? A.vbar_in[base + off] : zero2; };
    const float2 d1 = csub(r0, R[qx + (size_t)p.jm1 * M + p.i]), d2 = csub(r0, R[qx + (size_t)p.j * M + p.im1]);
    if (mode != 1) A.bbar[qx + o] = A.first ? r0 : cadd(A.bbar[qx + o], r0);
    if (mode == 2) {
      const float2 qa = csub(make_float2(2.f * rho * d1.x, 2.f * rho * d1.y), E(q1, o));
      const float2 qb = csub(make_float2(2.f * rho * d2.x, 2.f * rho * d2.y), E(q2, o));
      const float2 va = A.vck[q1 + o], vb = A.vck[q2 + o];
      A.ip_out[qx + o] = qa.x * va.x + qa.y * va.y + qb.x * vb.x + qb.y * vb.y;
    } else {
      const float2 d1n = csub(R[qx + oj], r0), d2n = csub(R[qx + oi], r0);   // channel 1 at (i,j+1), channel 2 at (i+1,j)
      float2 n1, n2, n1n, n2n;
      if (mode == 0) {
        n1 = bwd_point(d1, A.vck[q1 + o], E(q1, o), rho, tau, true, racc, tacc);
        n2 = bwd_point(d2, A.vck[q2 + o], E(q2, o), rho, tau, true, racc, tacc);
        n1n = bwd_point(d1n, A.vck[q1 + oj], E(q1, oj), rho, tau, false, racc, tacc);
        n2n = bwd_point(d2n, A.vck[q2 + oi], E(q2, oi), rho, tau, false, racc, tacc);
      } else {
        auto PIX = [&](size_t off, bool) { return A.sc[(size_t)grp * p.plane + off]; };   // (s, tau ip / n^3), k_iso_coef
        const float2 s0 = PIX(o, true), sj = PIX(oj, false), si = PIX(oi, false);
        n1 = iso_bwd_full(d1, A.vck[q1 + o], E(q1, o), rho, s0, true, racc);
        n2 = iso_bwd_full(d2, A.vck[q2 + o], E(q2, o), rho, s0, true, racc);
        n1n = iso_bwd_full(d1n, A.vck[q1 + oj], E(q1, oj), rho, sj, false, racc);
        n2n = iso_bwd_full(d2n, A.vck[q2 + oi], E(q2, oi), rho, si, false, racc);
      }
      A.vbar_out[q1 + o] = n1;
      A.vbar_out[q2 + o] = n2;
      xout[qx + o] = cadd(csub(n1, n1n), csub(n2, n2n));   // xbar_{k-1} = D^T vbar_{k-1}
    }
  }
  if (mode == 2) return;
  // scalar partial sums: one atomic pair per block when the whole block lies in one group, else (at most G-1 blocks) per thread
  const size_t per_group = (size_t)A.pm.Qg * M * A.N, first = (size_t)blockIdx.x * GK_NT;
  size_t last = first + GK_NT - 1;
  if (last >= per_group * A.pm.G) last = per_group * A.pm.G - 1;
  if (first / per_group == last / per_group) {
    const double rs = block_sum(racc), ts = block_sum(tacc);
    if (threadIdx.x == 0) {
      atomicAdd(A.acc + (size_t)A.AS * (int)(first / per_group) + acc_rho(A.ir), rs);
      atomicAdd(A.acc + (size_t)A.AS * (int)(first / per_group) + acc_tau(A.it), ts);
    }
  } else if (live) {
    if (racc != 0.0) atomicAdd(A.acc + (size_t)A.AS * grp + acc_rho(A.ir), racc);
    if (tacc != 0.0) atomicAdd(A.acc + (size_t)A.AS * grp + acc_tau(A.it), tacc);
  }
}

// ---- last backward step (k_dim1_bwd_last): bbar_total = bbar + rbar_1 -> spectrum input (mode 0) or ybar (mode 1) ----
__global__ void __launch_bounds__(GK_NT) gk_bwd_last(Dim1BwdArgs A, const float2* R, float2* xout, int M, int Q, int mode) {
  const size_t idx = (size_t)blockIdx.x * GK_NT + threadIdx.x;
  Pix p;
  if (!gk_pix(idx, M, A.N, Q, p)) return;
  float2 v = R[idx];
  if (!A.first) v = cadd(v, A.bbar[idx]);
  if (mode == 0) {
    xout[idx] = v;
    return;
  }
  const long ia = pm_in(A.pm, p.q, 0), ib = pm_in(A.pm, p.q, 1);
  if (A.pm.G > 1 && A.pm.in_gstride == 0) {
    atomicAdd(A.ybar + (size_t)ia * p.plane + p.o, v.x);
    if (ib >= 0) atomicAdd(A.ybar + (size_t)ib * p.plane + p.o, v.y);
  } else {
    A.ybar[(size_t)ia * p.plane + p.o] = v.x;
    if (ib >= 0) A.ybar[(size_t)ib * p.plane + p.o] = v.y;
  }
}

}  // namespace admmtv
