// args.cuh -- problem geometry, kernel argument blocks and the per-size launcher declarations.
// Each FFT length is compiled in its own translation unit (inst_dim1.cu / inst_dim2.cu with
// -DADMMTV_INST=<log2 length>) so the build parallelises; admmtv_api.cu only sees declarations.
#pragma once

#include "compat.cuh"

namespace admmtv {

// Plane bookkeeping for grouped calls: G independent problems of identical shape batched into one
// launch sequence (per-image PSFs / noise levels, or the 5 parallel branches of net_build.jl:113-128).
// Pairs never straddle groups: group g owns pairs [g*Qg, (g+1)*Qg) and Sg = P*Bg planes.
struct PlaneMap {
  int P, Sg, Qg, G;
  int in_gstride;  // planes between consecutive groups in the INPUT (Sg, or 0 when all groups share y)
  int concat;      // 1: output is (M,N,G*P,Bg), group g in channels [g*P,(g+1)*P)  (Flux chcat)
};
// index (in planes of N*M floats) of plane c (0/1) of pair q, or -1 if the pair's second plane is padding
ADMMTV_HD inline long pm_in(const PlaneMap& m, int q, int c) {
  const int g = q / m.Qg, sl = 2 * (q % m.Qg) + c;
  if (sl >= m.Sg) return -1;
  return (long)g * m.in_gstride + sl;
}
ADMMTV_HD inline long pm_out(const PlaneMap& m, int q, int c) {
  const int g = q / m.Qg, sl = 2 * (q % m.Qg) + c;
  if (sl >= m.Sg) return -1;
  if (!m.concat) return (long)g * m.Sg + sl;
  const int bl = sl / m.P, pp = sl % m.P;
  return (long)bl * (m.G * m.P) + g * m.P + pp;
}

struct Geom {
  int M, N, P, B, S, Q, LM, LN, K, kh, kw, nh;   // LM / LN: size ids (fft_core.cuh dim_id); 0 = generic-size kernels for that pass
  int planned;   // bit 0: dim-1 spectra in the plan's digit-reversed order (else natural); bit 1: same for dim 2
  int G, Bg, Sg, Qg;  // groups, images / planes / pairs per group (S = G*Sg planes, Q = G*Qg pairs)
  int iso;            // isotropic TV: the workspaces carry the per-pair partial sums of the per-pixel reductions
  int PS;             // lambda / rho values per group: 1 (the reference: one value for all iterations) or K (per-iteration
                      // learned parameters, ADMMTV_FLAG_PER_ITER_PARAMS); iteration k (1-based) uses entry PS == 1 ? 0 : k-1
  int AS;             // doubles per group in the scalar accumulator block `acc` (see acc_* below)
  int spatial;        // the backward differentiates the spatial H^T y path (kh > 0 and no NOGRAD_REPEAT): F y is checkpointed
  PlaneMap pm;
  size_t plane;  // N*M
  size_t pk;     // Q*N*M  (pair-packed complex elements)
};

// Scalar accumulator block of the backward, per group AS = 4 + 3 PS doubles:
//   [2] biasbar ; for parameter entry i: [4+3i] rhobar direct term, [5+3i] taubar, [6+3i] rhobar spectral term
ADMMTV_HD inline int acc_stride(int PS) { return 4 + 3 * PS; }
ADMMTV_HD inline int acc_bias() { return 2; }
ADMMTV_HD inline int acc_rho(int i) { return 4 + 3 * i; }
ADMMTV_HD inline int acc_tau(int i) { return 5 + 3 * i; }
ADMMTV_HD inline int acc_rhos(int i) { return 6 + 3 * i; }

struct PackArgs {
  const float* src;    // MODE 0/1: (M,N,S) planes
  const float2* src_packed;  // MODE 2: [Q][N][M] pair-packed spatial data
  float2* packed_out;  // MODE 0: optional pair-packed copy of the input (b = y when h is empty)
  const float* xout;   // MODE 1
  float2* spec;        // [Q][N][M]
  const float2* twM;
  const float* target; // MODE 3: the cotangent is formed on the fly, xbar = mse_scale (x_out - target)  (MSE loss pullback seed)
  float mse_scale;     //         2 / numel
  double* loss_acc;    //         += sum (x_out - target)^2
  double* bias_acc;    // MODE 1 / 3, may be null: biasbar slot of group 0 (groups are acc_stride apart)
  int acc_stride;      // doubles between the groups' accumulator blocks
  PlaneMap pm;
  int N, S, act;
};

struct OutArgs {
  const float2* spec;  // [Q][N][M]
  float2* packed;      // MODE 0
  float* planes;       // MODE 1: (M,N,S)
  const float* bias;   // MODE 1, may be null
  const float2* twM;
  PlaneMap pm;
  int N, S, act;
  float scale;         // MODE 0 / 2: factor applied to the result (1/M: the K tables carry only the dim-2 round trip's 1/N)
};

struct Dim1FwdArgs {
  const float2* spec_in;
  float2* spec_out;
  const float2* bpk;
  const float2* vprev;  // [Q][2][N][M]; ignored when !HAS_VPREV (v_0 = 0)
  float2* vnew;         // [Q][2][N][M]
  const float* nsq;     // isotropic: per-pixel shrink scale s [G][N][M] (pass B: s_k ; pass A: s_{k-1}), k_iso_scale
  float* nsq_out;       // isotropic pass A: [Q][N][M], pair q's share of |v_k|^2 per pixel (k_iso_scale adds them in order)
  const float2* twM;
  const float* lambda;  // [G][PS]
  const float* rho;     // [G][PS]
  int PS;               // parameter entries per group
  int ip, ic, in;       // entries of the previous iteration (tau of v_{k-1}), this one (tau of v_k) and the next (rho of r_{k+1})
  int N;
  int Qg;               // pairs per group: group of pair q = q / Qg
};

struct Dim2Args {
  const float2* in;    // [Q][N][M]
  float2* out;         // [Q][N][M]
  const float* ctab;   // MUL 0: real table [N][M]
  const float2* ktab;  // MUL 1/2: complex table [N][M] (2 = use its conjugate)
  float2* zsave;       // SAVE_Z: [Q][N][M] full spectrum before the multiply
  const float2* z2;    // ACC: second spectrum [Q][N][M]
  double* gacc;        // ACC 1: double [N][M] += Re(conj(Z) Z2): fp64, K x (pairs / block) partial sums meet here, one
                       //        atomic per element and block after the shared-memory pre-sum
  float2* pacc;        // ACC 2: float2 [N][M] += conj(Z) Z2: one launch, one term per pair and element; fp64 atomics made
                       //        this pass 210 -> 290 us without moving hbar (its floor is the fp32 rounding of the spectra)
  const float2* twN;
  int M;
  int Q;               // number of plane pairs (blocks loop q = blockIdx.y, += gridDim.y)
  int Qg;              // pairs per group
  size_t tab_stride;   // elements between the tables of consecutive groups (N*M, or 0 for one group)
};

struct Dim1BwdArgs {
  const float2* spec_in;   // dim-1 spectrum of rbar_k
  float2* spec_out;        // dim-1 spectrum of xbar_{k-1} (or of bbar_total for the last step)
  const float2* vck;       // v_{k-1}        [Q][2][N][M] (checkpoint)
  const float2* vbar_in;   // vbar_k         [Q][2][N][M] (ignored when !HAS_VBAR)
  float2* vbar_out;        // vbar_{k-1}
  float2* bbar;            // [Q][N][M] running sum of rbar_k
  float* ybar;             // last step, empty PSF: (M,N,S) planes
  const float2* twM;
  const float* lambda;     // [G][PS]
  const float* rho;
  int PS, ir, it;          // parameter entries per group; entry of rho_k (this step's x-update) and of tau_{k-1} (mask of v_{k-1})
  int AS;                  // accumulator doubles per group
  double* acc;             // scalar accumulator blocks (acc_* above)
  const float2* sc;        // isotropic pass B: per-pixel (s, tau ip / n^3) [G][N][M] (k_iso_coef)
  float* ip_out;           // isotropic pass A: [Q][N][M], pair q's share of <q, v_{k-1}> per pixel (k_iso_coef adds them in order)
  PlaneMap pm;
  int N, S;
  int first;               // 1: bbar is written, not accumulated (k = K)
};

// variants of k_dim2: (MUL, SAVE_Z, ACC, FWD_ONLY)
enum Dim2Variant {
  D2_C = 0,      // x C                                   (forward iteration, inference)
  D2_C_SAVE,     // save F r_k, then x C                  (forward iteration, training)
  D2_KCONJ,      // x conj(K)/N                           (b = H^T y)
  D2_C_ACCG,     // G += Re(conj(Z) Z2), then x C         (backward iteration)
  D2_FWDONLY,    // write the full 2-D spectrum, stop     (F y for the PSF-gradient correlation)
  D2_K_ACCP,     // P += conj(Z) Z2, then x K/N           (ybar = H bbar and hbar correlation)
  D2_K,          // x K/N                                 (ybar = H bbar only)
  D2_KCONJ_SAVE  // save F y, then x conj(K)/N            (b = H^T y, training: F y feeds the PSF-gradient correlation)
};

template <int LM>
struct Dim1Launch {
  static int pack_fft1(const Geom& g, int mode, const PackArgs& a, cudaStream_t st);
  static int out(const Geom& g, int mode, const OutArgs& a, cudaStream_t st);
  static int fwd(const Geom& g, bool has_vprev, const Dim1FwdArgs& a, cudaStream_t st);
  static int fwd_iso_b(const Geom& g, const Dim1FwdArgs& a, cudaStream_t st);
  static int fwd_iso_a(const Geom& g, bool has_vprev, const Dim1FwdArgs& a, cudaStream_t st);
  static int bwd_iso_b(const Geom& g, bool has_vbar, const Dim1BwdArgs& a, cudaStream_t st);
  static int bwd_iso_a(const Geom& g, bool has_vbar, const Dim1BwdArgs& a, cudaStream_t st);
  static int bwd(const Geom& g, bool has_vbar, const Dim1BwdArgs& a, cudaStream_t st);
  static int bwd_last(const Geom& g, int mode, const Dim1BwdArgs& a, cudaStream_t st);
  static int col_tile();   // output columns per block of the tuned kernels (N must be a multiple of it); 0 for the generic launcher
};
template <int LN>
struct Dim2Launch {
  static int run(const Geom& g, int variant, const Dim2Args& a, cudaStream_t st);
  static int row_tile();   // rows per block of the tuned kernel (M must be a multiple of it); 0 for the generic launcher
};

#define ADMMTV_CHECK_LAUNCH()                      \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

}  // namespace admmtv
