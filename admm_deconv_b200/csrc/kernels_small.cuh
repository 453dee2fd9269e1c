// kernels_small.cuh -- k_small: ALL unrolled iterations of the anisotropic forward for planes that fit one SM.
//
// A plane pair of M x N <= 128 x 128 complex values is 128 KB: it fits the shared memory of one SM.  The two-launch
// iteration (k_dim2 + k_dim1_fwd, kernels.cuh) then pays 2 K launches and streams the spectrum through HBM twice per
// iteration (36 B per plane-pixel-iteration) for nothing.  Here ONE persistent block owns a pair for the whole call:
//   X (shared, [N][M] float2) holds r_k -> dim-1 FFT -> dim-2 FFT -> x C (ops.jl:168) -> dim-2 IFFT -> dim-1 IFFT -> x_k,
//   then the stencil sweep (ops.jl:169-173: D x, shrinkage, dual ascent, r_{k+1} = H^T y + rho D^T(z - u)) rewrites X in
//   place with r_{k+1}.  Only the state v_k = D x_k + u_{k-1} (16 B per pair-pixel, read + written), H^T y (8 B, read)
//   and the C table (4 B, L2-resident unless every image has its own PSF) move through global memory: 20 B per
//   plane-pixel-iteration instead of 36, and the ping-pong state of a block stays in L2 between its iterations.
// This is the path of BASELINE configs[4] (1024 x 128 x 128, per-image PSFs): there every pair carries ONE real plane
// (pairs never straddle groups), so the two-launch path transforms a zero plane per pair and reached 0.36 of the
// roofline.  Inference only (no checkpoint), anisotropic only (the isotropic norm couples the pairs of a call).
//
// STR ("straddling pairs", per-image PSFs / parameters with single-plane groups): instead of pairing every image with a zero
// plane, images 2q and 2q+1 -- which have DIFFERENT C tables -- share one complex transform.  With Z = F(a + i b) the two
// spectra are A = (Z[k] + conj Z[-k]) / 2 and B = (Z[k] - conj Z[-k]) / 2i, so the spectral division becomes
//     X'[k] = C_a A + i C_b B = Z[k] (C_a + C_b)/2 + conj(Z[-k]) (C_a - C_b)/2 ,
// a pointwise operation on MIRRORED element pairs (k, -k) -- both are in this block's shared memory.  It costs one extra pass
// over the tile (the multiply can no longer be fused between the last forward and first inverse radix stage) and halves the
// number of transforms.  C is real and even (C[-k] = C[k]); tau and rho become per-plane (float2).
#pragma once

#include "kernels.cuh"

namespace admmtv {

struct SmallArgs {
  const float2* bpk;    // [Q][N][M] H^T y (pair-packed, spatial)
  float2 *v0, *v1;      // [Q][2][N][M] state ping-pong
  const float* ctab;    // [G][PS][N][M]
  const float2 *twM, *twN;
  const float* lambda;  // [G][PS]
  const float* rho;
  float* planes;        // output (M,N,S) user layout
  const float* bias;    // [G] or null
  PlaneMap pm;
  size_t tab_stride;    // floats between the C tables of consecutive groups (0 for one group)
  int K, PS, act, Q;
  int G;                // STR: number of single-plane groups (images); pair q = images 2q, 2q+1; Q = ceil(G / 2)
};

template <int LM, int LN>
struct SmallCfg {
  static constexpr int M = dim_len(LM), N = dim_len(LN);
  static constexpr int NT = (M * N >= 128 * 128) ? 512 : 256;
  static constexpr int RPT = 2;                     // rows per thread in the stencil sweep
  static constexpr int NRP = M / RPT;               // row groups
  static constexpr int NR = NT / NRP;               // column ranges
  static constexpr int CPR = N / NR;                // columns per range
  static constexpr int CHUNK = CPR % 4 == 0 ? 4 : (CPR % 2 == 0 ? 2 : 1);
  static constexpr size_t SMEM = (size_t)M * N * sizeof(float2);
  static constexpr bool OK = is_pow2(M) && is_pow2(N) && M >= 32 && N >= 32 && SMEM <= 128 * 1024 && NT % NRP == 0 && NR >= 1 &&
                             N % NR == 0 && CPR >= 1;
};

// one dim-2 (strided) radix pass over the whole plane in shared memory: thread = (row i, work item wi), lanes along rows
template <int LM, int LN, int S, bool INV, int NT>
ADMMTV_DI void small_dim2_stage(float2* X, const float2* __restrict__ tw, int tid) {
  constexpr int M = dim_len(LM), N = dim_len(LN);
  using St = Stage<N, S, kP2>;
  for (int item = tid; item < M * St::ITEMS; item += NT) {
    const int i = item % M, wi = item / M;
    float2 p[St::R], a[St::R];
    stage_twiddles<N, S, INV, kP2>(wi, tw, p);
    const int base = St::base(wi);
#pragma unroll
    for (int m = 0; m < St::R; ++m) a[m] = X[sidx<LM>(base + m * St::STRIDE, i)];
    if (INV) stage_inv<N, S, kP2>(a, p);
    else stage_fwd<N, S, kP2>(a, p);
#pragma unroll
    for (int m = 0; m < St::R; ++m) X[sidx<LM>(base + m * St::STRIDE, i)] = a[m];
  }
}
template <int LM, int LN, int S, int NT>
ADMMTV_DI void small_dim2_fwd_up(float2* X, const float2* __restrict__ tw, int tid) {
  if constexpr (S < plan_stages(dim_len(LN), kP2) - 1) {
    small_dim2_stage<LM, LN, S, false, NT>(X, tw, tid);
    __syncthreads();
    small_dim2_fwd_up<LM, LN, S + 1, NT>(X, tw, tid);
  }
}
template <int LM, int LN, int S, int NT>
ADMMTV_DI void small_dim2_inv_down(float2* X, const float2* __restrict__ tw, int tid) {
  if constexpr (S >= 0) {
    small_dim2_stage<LM, LN, S, true, NT>(X, tw, tid);
    __syncthreads();
    small_dim2_inv_down<LM, LN, S - 1, NT>(X, tw, tid);
  }
}

struct Shrunk2 {
  float2 u, w;
};
// per-plane thresholds: component .x belongs to plane a, .y to plane b
ADMMTV_DI Shrunk shrink2(float2 v, float2 tau) {
  const float2 z = make_float2(st1(v.x, tau.x), st1(v.y, tau.y));
  Shrunk s;
  s.u = csub(v, z);
  s.w = csub(z, s.u);
  return s;
}

template <int LM, int LN, bool STR>
__global__ void __launch_bounds__(SmallCfg<LM, LN>::NT, 1) k_small(SmallArgs A) {
  using Cfg = SmallCfg<LM, LN>;
  constexpr int M = Cfg::M, N = Cfg::N, NT = Cfg::NT, RPT = Cfg::RPT, CPR = Cfg::CPR, CHUNK = Cfg::CHUNK;
  constexpr int NSM = plan_stages(M), NSN = plan_stages(N, kP2);
  using StL = Stage<N, NSN - 1, kP2>;
  ADMMTV_DYN_SMEM(float2, X);   // [N][M], column c at X + c*M, swizzled rows (sidx)
  const int tid = threadIdx.x;
  const size_t plane = (size_t)M * N;
  const int rg = tid % Cfg::NRP, cr = tid / Cfg::NRP;   // stencil ownership: rows i0.., columns c0 .. c0 + CPR - 1
  const int i0 = rg * RPT, c0 = cr * CPR;
  __shared__ short mirM[STR ? M : 1], mirN[STR ? N : 1];   // STR: storage position of the negated frequency, per dimension
  if (STR) {
    for (int p = tid; p < M; p += NT) mirM[p] = (short)freq_to_pos(M, (M - pos_to_freq(M, p)) % M);
    for (int p = tid; p < N; p += NT) mirN[p] = (short)freq_to_pos(N, (N - pos_to_freq(N, p, true, kP2)) % N, kP2);
    __syncthreads();
  }

  for (int q = blockIdx.x; q < A.Q; q += gridDim.x) {
    // groups (parameter / table owners) of the pair's two planes
    const int grp = STR ? 2 * q : q / A.pm.Qg;
    const int grb = STR ? (2 * q + 1 < A.G ? 2 * q + 1 : 2 * q) : grp;
    const float2* bq = A.bpk + (size_t)q * plane;
    const float* ctg = A.ctab + (size_t)grp * A.tab_stride;
    const float* ctgb = A.ctab + (size_t)grb * A.tab_stride;
    // r_1 = b
    for (int e = tid; e < (int)plane / 2; e += NT) {
      const int c = (2 * e) / M, i = (2 * e) % M;
      const float4 v = *reinterpret_cast<const float4*>(bq + 2 * (size_t)e);
      X[sidx<LM>(c, i)] = make_float2(v.x, v.y);
      X[sidx<LM>(c, i + 1)] = make_float2(v.z, v.w);
    }
    __syncthreads();

    for (int k = 1; k <= A.K; ++k) {
      const int pe = A.PS == 1 ? 0 : k - 1;
      // ---- x_k = F^-1 C F r_k, entirely in shared memory ------------------------------------------------------
      dim1_fwd_stages_up<LM, NT, 0>(X, N, A.twM, tid);
      dim1_smem_stage<LM, NT, NSM - 1, false>(X, N, A.twM, tid);
      __syncthreads();
      small_dim2_fwd_up<LM, LN, 0, NT>(X, A.twN, tid);
      if (STR) {
        // complete the forward transform, then the spectral division on mirrored element pairs (see the header)
        small_dim2_stage<LM, LN, NSN - 1, false, NT>(X, A.twN, tid);
        __syncthreads();
        const float* cta = ctg + (size_t)pe * plane;
        const float* ctb = ctgb + (size_t)pe * plane;
        for (int e = tid; e < (int)plane; e += NT) {
          const int p1 = e % M, p2 = e / M;
          const int m1 = mirM[p1], m2 = mirN[p2];
          const int em = m2 * M + m1;
          if (em < e) continue;   // the pair is handled by its lower element
          const float ca = cta[e], cb = ctb[e];
          const float sgm = 0.5f * (ca + cb), dlt = 0.5f * (ca - cb);
          const float2 za = X[sidx<LM>(p2, p1)], zb = X[sidx<LM>(m2, m1)];
          X[sidx<LM>(p2, p1)] = make_float2(sgm * za.x + dlt * zb.x, sgm * za.y - dlt * zb.y);
          if (em != e) X[sidx<LM>(m2, m1)] = make_float2(sgm * zb.x + dlt * za.x, sgm * zb.y - dlt * za.y);
        }
        __syncthreads();
        small_dim2_stage<LM, LN, NSN - 1, true, NT>(X, A.twN, tid);
      } else {  // last forward dim-2 stage, x C, first inverse dim-2 stage
        const float* ct = ctg + (size_t)pe * plane;
        for (int item = tid; item < M * StL::ITEMS; item += NT) {
          const int i = item % M, wi = item / M;
          float2 a[StL::R];
          float cv[StL::R];
#pragma unroll
          for (int m = 0; m < StL::R; ++m) cv[m] = ct[(size_t)(wi * StL::R + m) * M + i];   // table loads in flight during the DFT
#pragma unroll
          for (int m = 0; m < StL::R; ++m) a[m] = X[sidx<LM>(wi * StL::R + m, i)];
          Dft<StL::R, false>::run(a);
#pragma unroll
          for (int m = 0; m < StL::R; ++m) a[m] = cscale(a[m], cv[m]);
          Dft<StL::R, true>::run(a);
#pragma unroll
          for (int m = 0; m < StL::R; ++m) X[sidx<LM>(wi * StL::R + m, i)] = a[m];
        }
      }
      __syncthreads();
      small_dim2_inv_down<LM, LN, NSN - 2, NT>(X, A.twN, tid);
      dim1_inv_stages_down<LM, NT, NSM - 1>(X, N, A.twM, tid);   // ends synced: X = x_k

      if (k == A.K) break;
      // ---- stencil sweep: v_k, shrinkage, dual ascent, r_{k+1} (in place) ---------------------------------------
      const bool has_prev = k > 1;
      const int pn = A.PS == 1 ? 0 : k, pp = (A.PS == 1 || k < 2) ? 0 : k - 2;
      auto par2 = [&](const float* t, int e) { return make_float2(t[grp * A.PS + e], t[grb * A.PS + e]); };   // (plane a, plane b)
      const float2 rho_n = par2(A.rho, pn);
      const float2 lc = par2(A.lambda, pe), rc = par2(A.rho, pe), lp = par2(A.lambda, pp), rp = par2(A.rho, pp);
      const float2 tau = make_float2(lc.x / rc.x, lc.y / rc.y), tau_p = make_float2(lp.x / rp.x, lp.y / rp.y);
      const float2* vp = ((k & 1) ? A.v1 : A.v0) + (size_t)q * 2 * plane;
      float2* vn = ((k & 1) ? A.v0 : A.v1) + (size_t)q * 2 * plane;
      const float2 *vp1 = vp, *vp2 = vp + plane;
      float2 *vn1 = vn, *vn2 = vn + plane;
      auto colw = [&](int c) { return c < 0 ? c + N : (c >= N ? c - N : c); };

      // boundary values this thread needs from columns other threads will overwrite: x at the column after its range,
      // and channel 1 (dim-2 difference) of its first column
      float2 xr[RPT], w1c[RPT];
      {
        const int cl = colw(c0 - 1), cn = colw(c0 + CPR);
        float2 up[RPT], vst[RPT];
        if (has_prev) load_rows<RPT>(vp1 + (size_t)c0 * M + i0, up);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          xr[r] = X[sidx<LM>(cn, i0 + r)];
          float2 v = csub(X[sidx<LM>(c0, i0 + r)], X[sidx<LM>(cl, i0 + r)]);
          if (has_prev) v = cadd(v, shrink2(up[r], tau_p).u);
          vst[r] = v;
          w1c[r] = shrink2(v, tau).w;
        }
        store_rows<RPT>(vn1 + (size_t)c0 * M + i0, vst);
      }
      __syncthreads();   // every boundary read happened before any r is written

      for (int cb = 0; cb < CPR; cb += CHUNK) {
        float2 g1[CHUNK][RPT], g2[CHUNK][RPT + 1], gb[CHUNK][RPT];
#pragma unroll
        for (int cc = 0; cc < CHUNK; ++cc) {   // every global load of the chunk is issued before any is consumed
          const int c = c0 + cb + cc, cn = colw(c + 1);
          if (has_prev) {
            load_rows<RPT>(vp1 + (size_t)cn * M + i0, g1[cc]);
            load_rows<RPT>(vp2 + (size_t)c * M + i0, g2[cc]);
            g2[cc][RPT] = vp2[(size_t)c * M + wrapm<M>(i0 + RPT)];
          }
          load_rows<RPT>(bq + (size_t)c * M + i0, gb[cc]);
        }
        float2 rr[CHUNK][RPT];
#pragma unroll
        for (int cc = 0; cc < CHUNK; ++cc) {
          const int c = c0 + cb + cc, cn = colw(c + 1);
          const bool last_own = cb + cc == CPR - 1;
          float2 xc[RPT + 2];
          xc[0] = X[sidx<LM>(c, wrapm<M>(i0 - 1))];
#pragma unroll
          for (int r = 0; r < RPT; ++r) xc[r + 1] = X[sidx<LM>(c, i0 + r)];
          xc[RPT + 1] = X[sidx<LM>(c, wrapm<M>(i0 + RPT))];
          float2 w1n[RPT], w2[RPT + 1];
          {  // channel 1 at column c+1 (stored by its owner: this thread unless c is the last column of the range)
            float2 vst[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
              const float2 xn = last_own ? xr[r] : X[sidx<LM>(cn, i0 + r)];
              float2 v = csub(xn, xc[r + 1]);
              if (has_prev) v = cadd(v, shrink2(g1[cc][r], tau_p).u);
              vst[r] = v;
              w1n[r] = shrink2(v, tau).w;
            }
            if (!last_own) store_rows<RPT>(vn1 + (size_t)cn * M + i0, vst);
          }
          {  // channel 2 at column c, rows i0 .. i0+RPT (the last one is the neighbour's, recomputed)
            float2 vst[RPT];
#pragma unroll
            for (int r = 0; r <= RPT; ++r) {
              float2 v = csub(xc[r + 1], xc[r]);
              if (has_prev) v = cadd(v, shrink2(g2[cc][r], tau_p).u);
              if (r < RPT) vst[r] = v;
              w2[r] = shrink2(v, tau).w;
            }
            store_rows<RPT>(vn2 + (size_t)c * M + i0, vst);
          }
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            const float2 dt = cadd(csub(w1c[r], w1n[r]), csub(w2[r], w2[r + 1]));   // D^T(z - u)
            rr[cc][r] = make_float2(gb[cc][r].x + rho_n.x * dt.x, gb[cc][r].y + rho_n.y * dt.y);
            w1c[r] = w1n[r];
          }
        }
        __syncthreads();   // every thread is done reading the x columns of this chunk
#pragma unroll
        for (int cc = 0; cc < CHUNK; ++cc)
#pragma unroll
          for (int r = 0; r < RPT; ++r) X[sidx<LM>(c0 + cb + cc, i0 + r)] = rr[cc][r];
      }
      __syncthreads();
    }

    // ---- x_K -> user layout, + bias, activation (ops.jl:175, deconv_admm.jl:222-224) ---------------------------------
    {
      const long ia = STR ? 2 * q : pm_out(A.pm, q, 0), ib = STR ? (2 * q + 1 < A.G ? 2 * q + 1 : -1) : pm_out(A.pm, q, 1);
      const bool has_b = ib >= 0;
      const float bias = A.bias ? A.bias[grp] : 0.f, biasb = A.bias ? A.bias[grb] : 0.f;
      float* pa = A.planes + (size_t)ia * plane;
      float* pb = A.planes + (size_t)(has_b ? ib : 0) * plane;
      for (int e = tid; e < (int)plane / 4; e += NT) {
        const int c = (4 * e) / M, i = (4 * e) % M;
        float2 v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = X[sidx<LM>(c, i + t)];
        const size_t off = (size_t)c * M + i;
        *reinterpret_cast<float4*>(pa + off) = make_float4(act_apply(v[0].x + bias, A.act), act_apply(v[1].x + bias, A.act),
                                                            act_apply(v[2].x + bias, A.act), act_apply(v[3].x + bias, A.act));
        if (has_b)
          *reinterpret_cast<float4*>(pb + off) = make_float4(act_apply(v[0].y + biasb, A.act), act_apply(v[1].y + biasb, A.act),
                                                              act_apply(v[2].y + biasb, A.act), act_apply(v[3].y + biasb, A.act));
      }
    }
    __syncthreads();   // X is reloaded for the block's next pair
  }
}

// STR: H^T y of the padded single-plane pairs ([G] pairs, .y = 0) -> straddling pairs ([ceil(G/2)], .x = image 2q, .y = image 2q+1)
static __global__ void __launch_bounds__(256) k_small_repack(const float2* __restrict__ padded, float2* __restrict__ out, size_t plane, int G) {
  const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
  const int q = blockIdx.y;
  if (e >= plane) return;
  const float a = padded[(size_t)(2 * q) * plane + e].x;
  const float b = 2 * q + 1 < G ? padded[(size_t)(2 * q + 1) * plane + e].x : 0.f;
  out[(size_t)q * plane + e] = make_float2(a, b);
}

// launcher (inst_small.cu): returns ADMMTV_ERR_UNSUPPORTED (-3) when there is no instantiation for (M, N)
int run_small(const Geom& g, const SmallArgs& a, bool straddle, cudaStream_t st);
bool small_supported(const Geom& g);

}  // namespace admmtv
