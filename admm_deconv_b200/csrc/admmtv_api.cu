// admmtv_api.cu -- the C ABI of include/admmtv.h: argument checking, workspace carving and the
// launch sequences.  No torch types, no global mutable state; every table lives in the caller's
// workspace.  Compiled by nvcc for sm_100a (product) -- see compat.cuh for the test-only
// emulation build.
#include "../../include/admmtv.h"
#include "args.cuh"
#include "setup_kernels.cuh"
#include "kernels_small.cuh"

#include <stdio.h>

namespace admmtv {

static inline size_t align256(size_t n) { return (n + 255) & ~size_t(255); }
static inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

static int dim2_row_tile(int LN);
static int dim1_col_tile(int LM);
static Geom geom(const admmtv_desc* d) {
  Geom g;
  g.M = d->M; g.N = d->N; g.P = d->P; g.B = d->B;
  g.S = d->P * d->B;
  g.Q = (g.S + 1) / 2;
  g.LM = dim_id(d->M); g.LN = dim_id(d->N);
  if (g.LM <= 0 || g.LN <= 0) {
    // No register-FFT plan for one of the lengths: that pass runs the generic kernels.  The other pass keeps its tuned
    // kernel when its tiling divides the unplanned length: the dim-2 kernel needs its row tile to divide M (16-byte
    // aligned row pairs) -- 720x1280, 1080x1920 frames with Julia (M,N) = (height, width) -- and the dim-1 kernels
    // need their column tile to divide N (1280x720).  Otherwise both passes are generic.
    const int tr = (g.LM <= 0 && g.LN > 0) ? dim2_row_tile(g.LN) : 0;
    const int co = (g.LN <= 0 && g.LM > 0) ? dim1_col_tile(g.LM) : 0;
    const int lm = g.LM, ln = g.LN;
    g.LM = g.LN = 0;
    if (tr > 0 && d->M % tr == 0) g.LN = ln;
    else if (co > 0 && d->N % co == 0) g.LM = lm;
  }
  g.planned = (g.LM > 0 ? 1 : 0) | (g.LN > 0 ? 2 : 0);
  g.K = d->iters; g.kh = d->kh; g.kw = d->kw; g.nh = d->kh * d->kw;
  g.G = d->groups > 1 ? d->groups : 1;
  g.iso = d->iso ? 1 : 0;
  g.PS = (d->flags & ADMMTV_FLAG_PER_ITER_PARAMS) ? d->iters : 1;
  g.AS = acc_stride(g.PS);
  g.spatial = (d->kh > 0 && !(d->flags & ADMMTV_FLAG_NOGRAD_REPEAT)) ? 1 : 0;
  g.Bg = d->B / g.G;
  g.Sg = d->P * g.Bg;
  g.Qg = (g.Sg + 1) / 2;
  g.Q = g.G * g.Qg;
  g.pm.P = d->P; g.pm.Sg = g.Sg; g.pm.Qg = g.Qg; g.pm.G = g.G;
  g.pm.in_gstride = (d->flags & ADMMTV_FLAG_SHARED_INPUT) ? 0 : g.Sg;
  g.pm.concat = (d->flags & ADMMTV_FLAG_CHANNEL_CONCAT) ? 1 : 0;
  g.plane = (size_t)d->N * d->M;
  g.pk = g.plane * g.Q;
  return g;
}

// ---- workspace / checkpoint carving ---------------------------------------------------------
struct Carver {
  unsigned char* base;
  size_t off;
  explicit Carver(void* p) : base(reinterpret_cast<unsigned char*>(p)), off(0) {}
  template <class T>
  T* take(size_t count) {
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align256(count * sizeof(T));
    return r;
  }
};

struct FwdWs {
  float2 *twM, *twN;
  double2* T;
  float* ctab;
  float2* ktab;
  float* mask;
  float2 *bpk, *specA, *specB, *v0, *v1;
  float* npart;        // isotropic: per-pair shares of |v_k|^2 per pixel [Q][N][M]
  float* nsq0;         // isotropic: per-pixel |v_k|^2 (inference; training keeps them in the checkpoint)
  float *isc0, *isc1;  // isotropic: per-pixel shrink scale ring
  size_t bytes;
};
static FwdWs carve_fwd(const Geom& g, void* ws) {
  Carver c(ws);
  FwdWs w;
  w.twM = c.take<float2>(g.M);
  w.twN = c.take<float2>(g.N);
  w.T = c.take<double2>((size_t)g.M * (g.kw > 0 ? g.kw : 1) * g.G);
  w.ctab = c.take<float>(g.plane * g.G * g.PS);
  w.ktab = c.take<float2>(g.plane * g.G);
  w.mask = c.take<float>((size_t)(g.nh + 2 * g.PS) * g.G);
  w.bpk = c.take<float2>(g.pk);
  w.specA = c.take<float2>(g.pk);
  w.specB = c.take<float2>(g.pk);
  w.v0 = c.take<float2>(2 * g.pk);
  w.v1 = c.take<float2>(2 * g.pk);
  const size_t ni = g.iso ? g.plane * g.G : 0;
  w.npart = c.take<float>(g.iso ? g.pk : 0);
  w.nsq0 = c.take<float>(ni);
  w.isc0 = c.take<float>(ni);
  w.isc1 = c.take<float>(ni);
  w.bytes = c.off;
  return w;
}

struct Ckpt {
  float* mask;   // [2 + nh]
  float2* vck;   // (K-1) slots of [Q][2][N][M]   : v_1 .. v_{K-1}
  float2* zck;   // K slots of [Q][N][M]          : F r_1 .. F r_K
  float* nck;    // isotropic: (K-1) slots of [N][M] : per-pixel |v_k|^2
  // what the forward computed once and the backward would otherwise recompute:
  float2 *twM, *twN;   // twiddles
  float* ctab;         // C / (M N)
  float2 *ktab, *sig;  // conj(K) / N and Sigma
  float2* yck;         // [Q][N][M] : F y (only when the backward differentiates the spatial H^T y path)
  size_t bytes;
};
static Ckpt carve_ckpt(const Geom& g, void* p) {
  Carver c(p);
  Ckpt k;
  k.mask = c.take<float>((size_t)(g.nh + 2 * g.PS) * g.G);
  k.vck = c.take<float2>((size_t)(g.K > 1 ? g.K - 1 : 0) * 2 * g.pk);
  k.zck = c.take<float2>((size_t)g.K * g.pk);
  k.nck = c.take<float>(g.iso ? (size_t)(g.K > 1 ? g.K - 1 : 0) * g.plane * g.G : 0);
  k.twM = c.take<float2>(g.M);
  k.twN = c.take<float2>(g.N);
  k.ctab = c.take<float>(g.plane * g.G * g.PS);
  k.ktab = c.take<float2>(g.kh > 0 ? g.plane * g.G : 0);
  k.sig = c.take<float2>(g.plane * g.G);
  k.yck = c.take<float2>(g.spatial ? g.pk : 0);
  k.bytes = c.off;
  return k;
}

#include "bwd_ws.inc"

// ---- per-size dispatch (definitions live in inst_dim1.cu / inst_dim2.cu) ------------------------
#define ADMMTV_SWITCH_LOG2(val, NAME, ...)                                   \
  switch (val) {                                                             \
    case 0: { constexpr int NAME = 0; __VA_ARGS__ } break;                   \
    case 5: { constexpr int NAME = 5; __VA_ARGS__ } break;                   \
    case 6: { constexpr int NAME = 6; __VA_ARGS__ } break;                   \
    case 7: { constexpr int NAME = 7; __VA_ARGS__ } break;                   \
    case 8: { constexpr int NAME = 8; __VA_ARGS__ } break;                   \
    case 9: { constexpr int NAME = 9; __VA_ARGS__ } break;                   \
    case 10: { constexpr int NAME = 10; __VA_ARGS__ } break;                 \
    case 11: { constexpr int NAME = 11; __VA_ARGS__ } break;                 \
    case 12: { constexpr int NAME = 12; __VA_ARGS__ } break;                 \
    case 20: { constexpr int NAME = 20; __VA_ARGS__ } break;                 \
    case 21: { constexpr int NAME = 21; __VA_ARGS__ } break;                 \
    case 22: { constexpr int NAME = 22; __VA_ARGS__ } break;                 \
    case 23: { constexpr int NAME = 23; __VA_ARGS__ } break;                 \
    case 24: { constexpr int NAME = 24; __VA_ARGS__ } break;                 \
    case 25: { constexpr int NAME = 25; __VA_ARGS__ } break;                 \
    case 26: { constexpr int NAME = 26; __VA_ARGS__ } break;                 \
    case 27: { constexpr int NAME = 27; __VA_ARGS__ } break;                 \
    case 28: { constexpr int NAME = 28; __VA_ARGS__ } break;                 \
    case 29: { constexpr int NAME = 29; __VA_ARGS__ } break;                 \
    case 30: { constexpr int NAME = 30; __VA_ARGS__ } break;                 \
    case 31: { constexpr int NAME = 31; __VA_ARGS__ } break;                 \
    default: break;                                                          \
  }                                                                          \
  return ADMMTV_ERR_UNSUPPORTED;

#define ADMMTV_SWITCH_LOG2_RC(val, NAME, RC, ...)                            \
  switch (val) {                                                             \
    case 0: { constexpr int NAME = 0; __VA_ARGS__ } break;                   \
    case 5: { constexpr int NAME = 5; __VA_ARGS__ } break;                   \
    case 6: { constexpr int NAME = 6; __VA_ARGS__ } break;                   \
    case 7: { constexpr int NAME = 7; __VA_ARGS__ } break;                   \
    case 8: { constexpr int NAME = 8; __VA_ARGS__ } break;                   \
    case 9: { constexpr int NAME = 9; __VA_ARGS__ } break;                   \
    case 10: { constexpr int NAME = 10; __VA_ARGS__ } break;                 \
    case 11: { constexpr int NAME = 11; __VA_ARGS__ } break;                 \
    case 12: { constexpr int NAME = 12; __VA_ARGS__ } break;                 \
    case 20: { constexpr int NAME = 20; __VA_ARGS__ } break;                 \
    case 21: { constexpr int NAME = 21; __VA_ARGS__ } break;                 \
    case 22: { constexpr int NAME = 22; __VA_ARGS__ } break;                 \
    case 23: { constexpr int NAME = 23; __VA_ARGS__ } break;                 \
    case 24: { constexpr int NAME = 24; __VA_ARGS__ } break;                 \
    case 25: { constexpr int NAME = 25; __VA_ARGS__ } break;                 \
    case 26: { constexpr int NAME = 26; __VA_ARGS__ } break;                 \
    case 27: { constexpr int NAME = 27; __VA_ARGS__ } break;                 \
    case 28: { constexpr int NAME = 28; __VA_ARGS__ } break;                 \
    case 29: { constexpr int NAME = 29; __VA_ARGS__ } break;                 \
    case 30: { constexpr int NAME = 30; __VA_ARGS__ } break;                 \
    case 31: { constexpr int NAME = 31; __VA_ARGS__ } break;                 \
    default: RC = ADMMTV_ERR_UNSUPPORTED; break;                             \
  }

static int run_pack_fft1(const Geom& g, int mode, const PackArgs& a, cudaStream_t st) {
  ADMMTV_SWITCH_LOG2(g.LM, LM, { return Dim1Launch<LM>::pack_fft1(g, mode, a, st); })
}
static int run_dim1_out(const Geom& g, int mode, const OutArgs& a, cudaStream_t st) {
  ADMMTV_SWITCH_LOG2(g.LM, LM, { return Dim1Launch<LM>::out(g, mode, a, st); })
}
static int run_dim1_fwd(const Geom& g, bool has_vprev, const Dim1FwdArgs& a, cudaStream_t st) {
  ADMMTV_SWITCH_LOG2(g.LM, LM, { return Dim1Launch<LM>::fwd(g, has_vprev, a, st); })
}
static int run_dim2(const Geom& g, Dim2Variant v, const Dim2Args& a, cudaStream_t st) {
  ADMMTV_SWITCH_LOG2(g.LN, LN, { return Dim2Launch<LN>::run(g, (int)v, a, st); })
}

static int dim2_row_tile(int LN) {
  ADMMTV_SWITCH_LOG2(LN, L, { return Dim2Launch<L>::row_tile(); })
}
static int dim1_col_tile(int LM) {
  ADMMTV_SWITCH_LOG2(LM, L, { return Dim1Launch<L>::col_tile(); })
}

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Enqueue the setup kernels shared by forward and backward.
static int run_setup(const Geom& g, const float* h, const float* rho, float2* twM, float2* twN, double2* T,
                     float* ctab, float2* ktab, float2* sig, cudaStream_t st) {
  {
    const int n = g.M > g.N ? g.M : g.N;
    ADMMTV_LAUNCH(k_setup_twiddles, dim3((n + 255) / 256), dim3(256), 0, st, twM, g.M, twN, g.N);
    ADMMTV_CHECK_LAUNCH();
  }
  // one table set per group (blockIdx.y = group)
  if (g.kh > 0) {
    const int n = g.M * g.kw;
    ADMMTV_LAUNCH(k_setup_psf_dim1, dim3((n + 127) / 128, (unsigned)g.G), dim3(128), 0, st, h, g.kh, g.kw, g.M, T);
    ADMMTV_CHECK_LAUNCH();
  }
  {
    const size_t n = g.plane;
    ADMMTV_LAUNCH(k_setup_tables, dim3((unsigned)((n + 127) / 128), (unsigned)g.G, (unsigned)g.PS), dim3(128), 0, st, (const double2*)T, g.kh,
                  g.kw, g.M, g.N, rho, ctab, g.kh > 0 ? ktab : (float2*)nullptr, sig, g.planned);
    ADMMTV_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace admmtv

using namespace admmtv;

extern "C" {

int admmtv_version(void) { return ADMMTV_VERSION; }

const char* admmtv_strerror(int code) {
  switch (code) {
    case ADMMTV_OK: return "ok";
    case ADMMTV_ERR_NULL: return "admmtv: null pointer argument";
    case ADMMTV_ERR_SHAPE: return "admmtv: invalid shape (dims must be positive, PSF no larger than the image)";
    case ADMMTV_ERR_UNSUPPORTED: return "admmtv: unsupported size (M and N must be at most 4096; SSIM windows at most 11 taps)";
    case ADMMTV_ERR_ITERS: return "admmtv: iters must be >= 1";
    case ADMMTV_ERR_ENUM: return "admmtv: invalid enum / flag value in descriptor";
    case ADMMTV_ERR_ALIGN: return "admmtv: workspace and checkpoint must be 256-byte aligned";
    case ADMMTV_ERR_NO_DEVICE: return "admmtv: CUDA device unavailable (there is no CPU fallback)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "admmtv: unknown error";
}

int admmtv_check(const admmtv_desc* d) {
  if (!d) return ADMMTV_ERR_NULL;
  if (d->M <= 0 || d->N <= 0 || d->P <= 0 || d->B <= 0) return ADMMTV_ERR_SHAPE;
  if (d->kh < 0 || d->kw < 0 || ((d->kh == 0) != (d->kw == 0))) return ADMMTV_ERR_SHAPE;
  if (d->kh > d->M || d->kw > d->N) return ADMMTV_ERR_SHAPE;
  if (dim_id(d->M) < 0 || dim_id(d->N) < 0) return ADMMTV_ERR_UNSUPPORTED;
  if (d->iters < 1) return ADMMTV_ERR_ITERS;
  if (d->iso != 0 && d->iso != 1) return ADMMTV_ERR_ENUM;
  if (d->activation < 0 || d->activation > 3) return ADMMTV_ERR_ENUM;
  if (d->has_bias != 0 && d->has_bias != 1) return ADMMTV_ERR_ENUM;
  if (d->flags & ~(ADMMTV_FLAG_NO_CLAMP | ADMMTV_FLAG_NOGRAD_REPEAT | ADMMTV_FLAG_SHARED_INPUT | ADMMTV_FLAG_CHANNEL_CONCAT |
                   ADMMTV_FLAG_ISO_PRECOMPUTE | ADMMTV_FLAG_ISO_INLINE | ADMMTV_FLAG_PER_ITER_PARAMS | ADMMTV_FLAG_NO_SMALL))
    return ADMMTV_ERR_ENUM;
  if (d->groups < 0 || (d->groups > 1 && d->B % d->groups != 0)) return ADMMTV_ERR_SHAPE;
  if (d->groups <= 1 && (d->flags & (ADMMTV_FLAG_SHARED_INPUT | ADMMTV_FLAG_CHANNEL_CONCAT))) return ADMMTV_ERR_ENUM;
  if (d->device < 0) return ADMMTV_ERR_ENUM;
  return ADMMTV_OK;
}

int admmtv_workspace_bytes(const admmtv_desc* d, size_t* fwd_bytes, size_t* ckpt_bytes, size_t* bwd_bytes) {
  int rc = admmtv_check(d);
  if (rc) return rc;
  const Geom g = geom(d);
  if (fwd_bytes) *fwd_bytes = carve_fwd(g, nullptr).bytes;
  if (ckpt_bytes) *ckpt_bytes = carve_ckpt(g, nullptr).bytes;
  if (bwd_bytes) *bwd_bytes = carve_bwd(g, nullptr).bytes;
  return ADMMTV_OK;
}

int admmtv_ckpt_layout(const admmtv_desc* d, size_t out[4]) {
  int rc = admmtv_check(d);
  if (rc) return rc;
  if (!out) return ADMMTV_ERR_NULL;
  const Geom g = geom(d);
  unsigned char* base = reinterpret_cast<unsigned char*>(uintptr_t(256));
  const Ckpt k = carve_ckpt(g, base);
  out[0] = (size_t)(reinterpret_cast<unsigned char*>(k.mask) - base);
  out[1] = (size_t)(reinterpret_cast<unsigned char*>(k.vck) - base);
  out[2] = (size_t)(reinterpret_cast<unsigned char*>(k.zck) - base);
  out[3] = (size_t)(reinterpret_cast<unsigned char*>(k.nck) - base);
  return ADMMTV_OK;
}

int admmtv_forward_launches(const admmtv_desc* d, int with_ckpt) {
  (void)with_ckpt;
  if (admmtv_check(d)) return 0;
  if (geom(d).LM == 0) {   // generic sizes: every dim-1 contract is [DFT] sweep [DFT] (inst_generic.cu)
    int n = 1 + 2 + (d->kh > 0 ? 1 : 0) + 2 /*pack, DFT*/ + (d->kh > 0 ? 1 + 2 + 2 : 0);
    n += d->iters + (d->iters - 1) * (d->iso ? 5 : 3) + 2;
    return n;
  }
  int n = 1 /*clamp*/ + 2 /*twiddles, tables*/ + (d->kh > 0 ? 1 : 0) + 1 /*pack*/ + (d->kh > 0 ? 3 : 0);
  n += d->iters + (d->iters - 1) * (d->iso ? 3 : 1) + 1;
  return n;
}

// Optional per-kernel-class CUDA-event timing (admmtv_profile_forward): class 0 = k_dim2 of the
// iterations, 1 = k_dim1_fwd, 2 = everything else (setup, H^T y, final).
struct Timing {
  enum { MAXL = 4096 };
  cudaEvent_t ev[MAXL + 1];
  int cls[MAXL];
  int n;
};
static void tm_mark(Timing* tm, cudaStream_t st, int cls_of_next) {
  if (!tm || tm->n >= Timing::MAXL) return;
  cudaEventRecord(tm->ev[tm->n], st);
  tm->cls[tm->n] = cls_of_next;
  tm->n++;
}

static int forward_impl(const admmtv_desc* d, const float* y, float* h, float* lambda, float* rho, const float* bias,
                        float* x_out, void* workspace, void* ckpt, void* stream, Timing* tm,
                        const admmtv_hooks* hk = nullptr) {
  int rc = admmtv_check(d);
  if (rc) return rc;
  if (!y || !lambda || !rho || !x_out || !workspace) return ADMMTV_ERR_NULL;
  if (d->kh > 0 && !h) return ADMMTV_ERR_NULL;
  if (d->has_bias && !bias) return ADMMTV_ERR_NULL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(ckpt) & 255)) return ADMMTV_ERR_ALIGN;
  DeviceGuard guard(d->device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const Geom g = geom(d);
  const FwdWs w = carve_fwd(g, workspace);
  Ckpt ck;
  if (ckpt) ck = carve_ckpt(g, ckpt);

  tm_mark(tm, st, 2);
  // deconv_admm.jl:216-219 (persisted clamp) + gradient masks for the pullback
  ADMMTV_LAUNCH(k_clamp_params, dim3((unsigned)g.G), dim3(128), 0, st, lambda, rho, h, g.nh, g.PS, d->creg,
                (d->flags & ADMMTV_FLAG_NO_CLAMP) ? 0 : 1, ckpt ? ck.mask : w.mask);
  ADMMTV_CHECK_LAUNCH();
  // training: twiddles and tables go to the checkpoint, where the backward finds them (no second setup)
  const float2 *twM = ckpt ? ck.twM : w.twM, *twN = ckpt ? ck.twN : w.twN;
  const float* ctab = ckpt ? ck.ctab : w.ctab;
  const float2* ktab = ckpt ? ck.ktab : w.ktab;
  if ((rc = run_setup(g, h, rho, const_cast<float2*>(twM), const_cast<float2*>(twN), w.T, const_cast<float*>(ctab),
                      const_cast<float2*>(ktab), ckpt ? ck.sig : nullptr, st)))
    return rc;

  // Planes that fit one SM (kernels_small.cuh): every iteration inside ONE persistent kernel, the spectrum never leaves
  // shared memory.  Inference, anisotropic, enough pairs to occupy the GPU; ADMMTV_FLAG_NO_SMALL keeps the two-launch path.
  const bool use_small = !ckpt && !d->iso && !hk && !(d->flags & ADMMTV_FLAG_NO_SMALL) && small_supported(g) && g.Q >= 64;

  // spectrum ping-pong: `cur` holds the dim-1 spectrum of r_k, `oth` receives the k_dim2 result
  float2 *cur = w.specA, *oth = w.specB;
  // y -> pair-pack -> dim-1 spectrum (ops.jl:101 + first FFT pass)
  {
    PackArgs a{};
    a.src = y; a.spec = w.specA; a.twM = twM; a.N = g.N; a.S = g.S; a.pm = g.pm;
    a.packed_out = g.kh > 0 ? nullptr : w.bpk;  // empty h: b = y (ops.jl:149-151)
    if ((rc = run_pack_fft1(g, 0, a, st))) return rc;
  }
  if (g.kh > 0) {
    // b = H^T y = F^-1( conj(K) F y )  (ops.jl:163,168; hoisted out of the loop)
    Dim2Args a{};
    a.in = w.specA; a.out = w.specB; a.ktab = ktab; a.twN = twN; a.M = g.M;
    a.Qg = g.Qg; a.tab_stride = g.G > 1 ? g.plane : 0;
    const bool save_fy = ckpt && g.spatial;   // F y is the second factor of the PSF-gradient correlation
    if (save_fy) a.zsave = ck.yck;
    if ((rc = run_dim2(g, save_fy ? D2_KCONJ_SAVE : D2_KCONJ, a, st))) return rc;
    OutArgs o{};
    o.spec = w.specB; o.packed = w.bpk; o.twM = twM; o.N = g.N; o.S = g.S; o.pm = g.pm; o.scale = 1.f / (float)g.M;
    if ((rc = run_dim1_out(g, 0, o, st))) return rc;
    // r_1 = b: its dim-1 spectrum is taken from the ROUNDED spatial b (one more dim-1 pass) rather than from specB, which
    // already holds it before rounding.  Reusing specB saves a launch (104 us at 64 x 512^2 x 3) but makes r_1 inconsistent
    // with the r_k = b + rho D^T(...) of the later iterations at the rounding level, and the teacher-forced parameter
    // gradients then sit 1.7x further from the fp64 adjoint (hbar 4.4e-6 -> 1.25e-5 on 2 x 256^2, K = 10): not worth 1 %.
    if (!use_small) {
      PackArgs p{};
      p.src_packed = w.bpk; p.spec = w.specA; p.twM = twM; p.N = g.N; p.S = g.S; p.pm = g.pm;
      if ((rc = run_pack_fft1(g, 2, p, st))) return rc;
    }
  }

  if (use_small) {
    SmallArgs s{};
    s.bpk = w.bpk; s.v0 = w.v0; s.v1 = w.v1; s.ctab = ctab; s.twM = twM; s.twN = twN; s.lambda = lambda; s.rho = rho;
    s.planes = x_out; s.bias = d->has_bias ? bias : nullptr; s.pm = g.pm; s.tab_stride = g.G > 1 ? g.plane * g.PS : 0;
    s.K = g.K; s.PS = g.PS; s.act = d->activation; s.Q = g.Q;
    tm_mark(tm, st, 0);
    // per-image PSFs / parameters on single-plane images (groups of one plane): two images share one complex transform
    const bool straddle = g.G > 1 && g.Sg == 1 && !g.pm.concat && g.pm.in_gstride != 0;
    if ((rc = run_small(g, s, straddle, st))) return rc;
    tm_mark(tm, st, -1);
    return ADMMTV_OK;
  }

  // the unrolled iterations (ops.jl:166-174)
  for (int k = 1; k <= g.K; ++k) {
    Dim2Args a{};
    const int pe = g.PS == 1 ? 0 : k - 1;   // parameter entry of iteration k
    a.in = cur; a.out = oth; a.ctab = ctab + (size_t)pe * g.plane; a.twN = twN; a.M = g.M;
    a.Qg = g.Qg; a.tab_stride = g.G > 1 ? g.plane * g.PS : 0;
    if (ckpt) a.zsave = ck.zck + (size_t)(k - 1) * g.pk;
    tm_mark(tm, st, 0);
    if ((rc = run_dim2(g, ckpt ? D2_C_SAVE : D2_C, a, st))) return rc;
    if (k < g.K && !d->iso) {
      Dim1FwdArgs f{};
      f.spec_in = oth; f.spec_out = cur; f.bpk = w.bpk; f.twM = twM;
      f.lambda = lambda; f.rho = rho; f.N = g.N; f.Qg = g.Qg;
      f.PS = g.PS; f.ic = pe; f.ip = (g.PS == 1 || k < 2) ? 0 : k - 2; f.in = g.PS == 1 ? 0 : k;   // k < K here: entry k exists
      if (ckpt) {
        f.vprev = k > 1 ? ck.vck + (size_t)(k - 2) * 2 * g.pk : nullptr;
        f.vnew = ck.vck + (size_t)(k - 1) * 2 * g.pk;
      } else {
        f.vprev = (k & 1) ? w.v1 : w.v0;
        f.vnew = (k & 1) ? w.v0 : w.v1;
      }
      tm_mark(tm, st, 1);
      if ((rc = run_dim1_fwd(g, k > 1, f, st))) return rc;
    } else if (k < g.K) {
      // isotropic: v_k and the per-pixel norm first (pass A), then shrink + D^T + FFT (pass B)
      const float2* v_in;
      float2* v_out;
      float* nsq_new;                                    // |v_k|^2 per pixel (checkpointed for the backward)
      const float* s_prev = (k & 1) ? w.isc1 : w.isc0;   // s_{k-1}
      float* s_new = (k & 1) ? w.isc0 : w.isc1;          // s_k
      if (ckpt) {
        v_in = k > 1 ? ck.vck + (size_t)(k - 2) * 2 * g.pk : nullptr;
        v_out = ck.vck + (size_t)(k - 1) * 2 * g.pk;
        nsq_new = ck.nck + (size_t)(k - 1) * g.plane * g.G;
      } else {
        v_in = (k & 1) ? w.v1 : w.v0;
        v_out = (k & 1) ? w.v0 : w.v1;
        nsq_new = w.nsq0;
      }
      tm_mark(tm, st, 1);
      Dim1FwdArgs fa{};
      fa.spec_in = oth; fa.twM = twM; fa.lambda = lambda; fa.rho = rho; fa.N = g.N; fa.Qg = g.Qg;
      fa.PS = g.PS; fa.ic = pe; fa.ip = (g.PS == 1 || k < 2) ? 0 : k - 2; fa.in = g.PS == 1 ? 0 : k;
      fa.vprev = v_in; fa.vnew = v_out; fa.nsq = s_prev; fa.nsq_out = w.npart;
      ADMMTV_SWITCH_LOG2_RC(g.LM, LM, rc, { rc = Dim1Launch<LM>::fwd_iso_a(g, k > 1, fa, st); })
      if (rc) return rc;
      // the pairs' shares -> |v_k|^2 per pixel in a fixed order (bit-reproducible), then the shrink scale s_k
      const dim3 sgrid((unsigned)((g.plane + 255) / 256), (unsigned)g.G);
      const bool xrank = hk && hk->allreduce_sum;   // exact global-batch norm across ranks (admmtv_forward_ex)
      ADMMTV_LAUNCH(k_iso_scale, sgrid, dim3(256), 0, st, (const float*)w.npart, g.Qg, nsq_new, (const float*)lambda,
                    (const float*)rho, g.PS, pe, xrank ? (float*)nullptr : s_new, (int)g.plane);
      ADMMTV_CHECK_LAUNCH();
      if (xrank) {
        if ((rc = hk->allreduce_sum(nsq_new, g.plane * g.G, stream, hk->user))) return rc;
        ADMMTV_LAUNCH(k_iso_scale, sgrid, dim3(256), 0, st, (const float*)nullptr, g.Qg, nsq_new, (const float*)lambda,
                      (const float*)rho, g.PS, pe, s_new, (int)g.plane);
        ADMMTV_CHECK_LAUNCH();
      }
      Dim1FwdArgs f{};
      f.spec_out = cur; f.bpk = w.bpk; f.twM = twM; f.lambda = lambda; f.rho = rho; f.N = g.N; f.Qg = g.Qg;
      f.PS = g.PS; f.ic = pe; f.ip = fa.ip; f.in = fa.in;
      f.vprev = v_out; f.nsq = s_new;
      ADMMTV_SWITCH_LOG2_RC(g.LM, LM, rc, { rc = Dim1Launch<LM>::fwd_iso_b(g, f, st); })
      if (rc) return rc;
    }
  }
  tm_mark(tm, st, 2);
  // x_K -> user layout, + bias, activation (ops.jl:175, deconv_admm.jl:222-224)
  {
    OutArgs o{};
    o.spec = oth; o.planes = x_out; o.bias = d->has_bias ? bias : nullptr; o.twM = twM;
    o.N = g.N; o.S = g.S; o.act = d->activation; o.pm = g.pm;
    if ((rc = run_dim1_out(g, 1, o, st))) return rc;
  }
  tm_mark(tm, st, -1);
  return ADMMTV_OK;
}

int admmtv_forward(const admmtv_desc* d, const float* y, float* h, float* lambda, float* rho, const float* bias,
                   float* x_out, void* workspace, void* ckpt, void* stream) {
  return forward_impl(d, y, h, lambda, rho, bias, x_out, workspace, ckpt, stream, nullptr);
}

int admmtv_forward_ex(const admmtv_desc* d, const float* y, float* h, float* lambda, float* rho, const float* bias,
                      float* x_out, void* workspace, void* ckpt, void* stream, const admmtv_hooks* hooks) {
  return forward_impl(d, y, h, lambda, rho, bias, x_out, workspace, ckpt, stream, nullptr, hooks);
}

int admmtv_profile_forward(const admmtv_desc* d, const float* y, float* h, float* lambda, float* rho, const float* bias,
                           float* x_out, void* workspace, void* ckpt, void* stream, float* ms_out) {
  if (!ms_out) return ADMMTV_ERR_NULL;
  int rc = admmtv_check(d);
  if (rc) return rc;
  if (2 * d->iters + 8 > Timing::MAXL) return ADMMTV_ERR_ITERS;
  DeviceGuard guard(d->device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  Timing* tm = new Timing();
  tm->n = 0;
  const int nev = 2 * d->iters + 8;
  for (int i = 0; i < nev; ++i) cudaEventCreate(&tm->ev[i]);
  rc = forward_impl(d, y, h, lambda, rho, bias, x_out, workspace, ckpt, stream, tm);
  cudaError_t e = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0 && e != cudaSuccess) rc = (int)e;
  ms_out[0] = ms_out[1] = ms_out[2] = ms_out[3] = 0.f;
  if (rc == 0) {
    for (int i = 0; i + 1 < tm->n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tm->ev[i], tm->ev[i + 1]);
      const int c = tm->cls[i];
      if (c >= 0 && c < 3) ms_out[1 + c] += ms;
      ms_out[0] += ms;
    }
  }
  for (int i = 0; i < nev; ++i) cudaEventDestroy(tm->ev[i]);
  delete tm;
  return rc;
}

#include "bwd_api.inc"

int admmtv_forward_host(const admmtv_desc* d, const float* y, float* h, float* lambda, float* rho, const float* bias,
                        float* x_out) {
  int rc = admmtv_check(d);
  if (rc) return rc;
  if (!y || !lambda || !rho || !x_out) return ADMMTV_ERR_NULL;
  DeviceGuard guard(d->device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  const Geom g = geom(d);
  size_t fwd = 0;
  admmtv_workspace_bytes(d, &fwd, nullptr, nullptr);
  const size_t nimg = g.plane * g.S * sizeof(float);
  if (d->groups > 1) return ADMMTV_ERR_UNSUPPORTED;  // host convenience path: single call only
  unsigned char* dev = nullptr;
  const size_t o_y = 0, o_x = align256(nimg), o_h = o_x + align256(nimg), o_l = o_h + align256((g.nh + 1) * sizeof(float)),
               o_r = o_l + 256, o_b = o_r + 256, o_ws = o_b + 256;
  cudaError_t e = cudaMalloc((void**)&dev, o_ws + fwd);
  if (e != cudaSuccess) return (int)e;
  cudaStream_t st = 0;
  cudaMemcpyAsync(dev + o_y, y, nimg, cudaMemcpyHostToDevice, st);
  if (g.nh) cudaMemcpyAsync(dev + o_h, h, g.nh * sizeof(float), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dev + o_l, lambda, sizeof(float), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dev + o_r, rho, sizeof(float), cudaMemcpyHostToDevice, st);
  if (d->has_bias && bias) cudaMemcpyAsync(dev + o_b, bias, sizeof(float), cudaMemcpyHostToDevice, st);
  rc = admmtv_forward(d, (const float*)(dev + o_y), g.nh ? (float*)(dev + o_h) : nullptr, (float*)(dev + o_l),
                      (float*)(dev + o_r), d->has_bias ? (const float*)(dev + o_b) : nullptr, (float*)(dev + o_x),
                      dev + o_ws, nullptr, st);
  if (rc == 0) {
    cudaMemcpyAsync(x_out, dev + o_x, nimg, cudaMemcpyDeviceToHost, st);
    if (g.nh) cudaMemcpyAsync(h, dev + o_h, g.nh * sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(lambda, dev + o_l, sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(rho, dev + o_r, sizeof(float), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = (int)e;
  }
  cudaFree(dev);
  return rc;
}

}  // extern "C"
