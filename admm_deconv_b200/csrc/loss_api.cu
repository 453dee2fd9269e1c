// loss_api.cu -- the C ABI of include/admmtv_loss.h: argument checks, workspace carving, launches.
#include "../../include/admmtv.h"
#include "../../include/admmtv_batch.h"
#include "../../include/admmtv_loss.h"
#include "loss_kernels.cuh"

namespace admmtv {
namespace {

inline size_t up256(size_t n) { return (n + 255) & ~size_t(255); }

struct DevGuard {
  int prev;
  bool ok;
  explicit DevGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int check_shape(int M, int N, int C, int B) {
  if (M <= 0 || N <= 0 || C <= 0 || B <= 0) return ADMMTV_ERR_SHAPE;
  return ADMMTV_OK;
}

// sigma = 1.5, length 11 (src/metrics/ssim.jl:6-17)
const float kSsimGauss[11] = {0.00102838008447911f, 0.007598758135239185f, 0.03600077212843083f, 0.10936068950970002f,
                              0.2130055377112537f,  0.26601172486179436f,  0.2130055377112537f,  0.10936068950970002f,
                              0.03600077212843083f, 0.007598758135239185f, 0.00102838008447911f};

int ssim_args(SsimArgs& a, int M, int N, int C, int B, const float* taps, int L) {
  int rc = check_shape(M, N, C, B);
  if (rc) return rc;
  if (!taps) {
    taps = kSsimGauss;
    L = 11;
  }
  if (L < 1 || L > SS_LMAX) return ADMMTV_ERR_UNSUPPORTED;
  if (L > M || L > N) return ADMMTV_ERR_SHAPE;  // crop=true: the valid-size map would be empty
  a.M = M; a.N = N; a.C = C; a.B = B; a.L = L;
  a.Mo = M - L + 1; a.No = N - L + 1;
  for (int k = 0; k < SS_LMAX; ++k) a.f[k] = k < L ? taps[L - 1 - k] : 0.f;
  return ADMMTV_OK;
}

// general window: R separable terms u_r (L1 taps, dim 1) x v_r (L2 taps, dim 2), row-major [R][L]
int ssim_gen_args(SsimGenArgs& g, int M, int N, int C, int B, const float* u, const float* v, int L1, int L2, int R) {
  int rc = check_shape(M, N, C, B);
  if (rc) return rc;
  if (L1 < 1 || L1 > SS_LMAX || L2 < 1 || L2 > SS_LMAX || R < 1 || R > SS_LMAX) return ADMMTV_ERR_UNSUPPORTED;
  if (L1 > M || L2 > N) return ADMMTV_ERR_SHAPE;  // the valid-size map would be empty
  SsimArgs& a = g.a;
  a.M = M; a.N = N; a.C = C; a.B = B; a.L = 0;
  a.Mo = M - L1 + 1; a.No = N - L2 + 1;
  g.L1 = L1; g.L2 = L2; g.R = R;
  for (int r = 0; r < R; ++r) {   // flipped: NNlib conv is a true convolution (ssim.jl:112-119)
    for (int k = 0; k < SS_LMAX; ++k) {
      g.fu[r * SS_LMAX + k] = (u && k < L1) ? u[r * L1 + (L1 - 1 - k)] : 0.f;
      g.fv[r * SS_LMAX + k] = (v && k < L2) ? v[r * L2 + (L2 - 1 - k)] : 0.f;
    }
  }
  return ADMMTV_OK;
}

int pad_args(PadArgs& a, int M, int N, int planes, int lo1, int hi1, int lo2, int hi2) {
  if (M <= 0 || N <= 0 || planes <= 0) return ADMMTV_ERR_SHAPE;
  if (lo1 < 0 || hi1 < 0 || lo2 < 0 || hi2 < 0) return ADMMTV_ERR_SHAPE;
  if (lo1 > M || hi1 > M || lo2 > N || hi2 > N) return ADMMTV_ERR_SHAPE;   // NNlib: the pad cannot exceed the array size
  a.M = M; a.N = N; a.planes = planes; a.lo1 = lo1; a.hi1 = hi1; a.lo2 = lo2; a.hi2 = hi2;
  return ADMMTV_OK;
}

unsigned pad_grid(size_t total) {
  size_t nb = (total + 255) / 256;
  return (unsigned)(nb > 148 * 64 ? 148 * 64 : (nb ? nb : 1));
}

}  // namespace
}  // namespace admmtv

using namespace admmtv;

extern "C" {

int admmtv_gmsd_workspace_bytes(int M, int N, int C, int B, size_t* bytes) {
  int rc = check_shape(M, N, C, B);
  if (rc) return rc;
  if (bytes) *bytes = 2 * up256(2 * (size_t)B * sizeof(double));
  return ADMMTV_OK;
}

int admmtv_gmsd_forward(int M, int N, int C, int B, int device, const float* x, const float* y, float t, float alpha,
                        float* loss_out, void* workspace, void* stream) {
  int rc = check_shape(M, N, C, B);
  if (rc) return rc;
  if (!x || !y || !loss_out || !workspace) return ADMMTV_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return ADMMTV_ERR_ALIGN;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GmsdArgs a{};
  a.x = x; a.y = y; a.M = M; a.N = N; a.C = C; a.B = B; a.t = t; a.alpha = alpha;
  a.acc = reinterpret_cast<double*>(workspace);
  a.stats = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + up256(2 * (size_t)B * sizeof(double)));
  a.out = loss_out;
  cudaError_t e = cudaMemsetAsync(a.acc, 0, 2 * (size_t)B * sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  a.tiles_i = (M + GS_RF - 1) / GS_RF; a.tiles_j = (N + GS_CW - 1) / GS_CW;
  const size_t nwarp = (size_t)a.tiles_i * a.tiles_j * C * B, nblk_s = (nwarp + GS_NT / 32 - 1) / (GS_NT / 32);
  if (nblk_s > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  ADMMTV_LAUNCH(k_gmsd_fwd_s, dim3((unsigned)nblk_s), dim3(GS_NT), 0, st, a);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  ADMMTV_LAUNCH(k_gmsd_finalize, dim3(1), dim3(128), 0, st, a);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  return ADMMTV_OK;
}

int admmtv_gmsd_backward(int M, int N, int C, int B, int device, const float* x, const float* y, float t, float alpha,
                         const float* lossbar, const void* workspace, float* xbar, void* stream) {
  int rc = check_shape(M, N, C, B);
  if (rc) return rc;
  if (!x || !y || !lossbar || !workspace || !xbar) return ADMMTV_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return ADMMTV_ERR_ALIGN;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GmsdArgs a{};
  a.x = x; a.y = y; a.M = M; a.N = N; a.C = C; a.B = B; a.t = t; a.alpha = alpha;
  unsigned char* ws = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(workspace));
  a.acc = reinterpret_cast<double*>(ws);
  a.stats = reinterpret_cast<double*>(ws + up256(2 * (size_t)B * sizeof(double)));
  a.lossbar = lossbar;
  a.out = xbar;
  a.tiles_i = (M + GS_RB - 1) / GS_RB; a.tiles_j = (N + GS_CW - 1) / GS_CW;
  const size_t nwarp = (size_t)a.tiles_i * a.tiles_j * C * B, nblk_s = (nwarp + GS_NT / 32 - 1) / (GS_NT / 32);
  if (nblk_s > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  ADMMTV_LAUNCH(k_gmsd_bwd_s, dim3((unsigned)nblk_s), dim3(GS_NT), 0, st, a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADMMTV_OK : (int)e;
}

int admmtv_ssim_workspace_bytes(int M, int N, int C, int B, int L, int with_grad, size_t* bytes) {
  SsimArgs a{};
  const float dummy[SS_LMAX] = {0};
  int rc = ssim_args(a, M, N, C, B, L > 0 ? dummy : nullptr, L);
  if (rc) return rc;
  size_t n = up256(sizeof(double));
  if (with_grad) n += up256(3 * (size_t)a.Mo * a.No * C * B * sizeof(float));
  if (bytes) *bytes = n;
  return ADMMTV_OK;
}

int admmtv_ssim_forward(int M, int N, int C, int B, int device, const float* x, const float* y, const float* taps, int L,
                        float peakval, int as_loss, float* out, void* workspace, int with_grad, void* stream) {
  SsimArgs a{};
  int rc = ssim_args(a, M, N, C, B, taps, L);
  if (rc) return rc;
  if (!x || !y || !out || !workspace) return ADMMTV_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return ADMMTV_ERR_ALIGN;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  a.x = x; a.y = y;
  a.C1 = (peakval * 0.01f) * (peakval * 0.01f);  // ssim.jl:101-102
  a.C2 = (peakval * 0.03f) * (peakval * 0.03f);
  a.tiles_i = (a.Mo + SS_T - 1) / SS_T; a.tiles_j = (a.No + SS_T - 1) / SS_T;
  a.acc = reinterpret_cast<double*>(workspace);
  a.maps = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + up256(sizeof(double)));
  a.with_grad = with_grad ? 1 : 0;
  a.as_loss = as_loss ? 1 : 0;
  a.out = out;
  cudaError_t e = cudaMemsetAsync(a.acc, 0, sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  const size_t nblk = (size_t)a.tiles_i * a.tiles_j * C * B;
  if (nblk > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  if (a.L == 11) ADMMTV_LAUNCH(k_ssim_fwd4<11>, dim3((unsigned)nblk), dim3(SS_NT), 0, st, a);
  else if (a.L == 5) ADMMTV_LAUNCH(k_ssim_fwd4<5>, dim3((unsigned)nblk), dim3(SS_NT), 0, st, a);
  else ADMMTV_LAUNCH(k_ssim_fwd<0>, dim3((unsigned)nblk), dim3(SS_NT), 0, st, a);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  ADMMTV_LAUNCH(k_ssim_finalize, dim3(1), dim3(1), 0, st, a);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  return ADMMTV_OK;
}

int admmtv_ssim_backward(int M, int N, int C, int B, int device, const float* x, const float* y, const float* taps, int L,
                         int as_loss, const float* outbar, const void* workspace, float* xbar, void* stream) {
  SsimArgs a{};
  int rc = ssim_args(a, M, N, C, B, taps, L);
  if (rc) return rc;
  if (!x || !y || !outbar || !workspace || !xbar) return ADMMTV_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return ADMMTV_ERR_ALIGN;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  a.x = x; a.y = y;
  a.tiles_i = (M + SS_T - 1) / SS_T; a.tiles_j = (N + SS_T - 1) / SS_T;
  unsigned char* ws = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(workspace));
  a.acc = reinterpret_cast<double*>(ws);
  a.maps = reinterpret_cast<float*>(ws + up256(sizeof(double)));
  a.with_grad = 1;
  a.as_loss = as_loss ? 1 : 0;
  a.outbar = outbar;
  a.out = xbar;
  const size_t nblk = (size_t)a.tiles_i * a.tiles_j * C * B;
  if (nblk > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  if (a.L == 11) ADMMTV_LAUNCH(k_ssim_bwd4<11>, dim3((unsigned)nblk), dim3(SS_NT), 0, st, a);
  else if (a.L == 5) ADMMTV_LAUNCH(k_ssim_bwd4<5>, dim3((unsigned)nblk), dim3(SS_NT), 0, st, a);
  else ADMMTV_LAUNCH(k_ssim_bwd<0>, dim3((unsigned)nblk), dim3(SS_NT), 0, st, a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADMMTV_OK : (int)e;
}

int admmtv_ssim_window_workspace_bytes(int M, int N, int C, int B, int L1, int L2, int with_grad, size_t* bytes) {
  SsimGenArgs g{};
  int rc = ssim_gen_args(g, M, N, C, B, nullptr, nullptr, L1, L2, 1);
  if (rc) return rc;
  size_t n = up256(sizeof(double));
  if (with_grad) n += up256(3 * (size_t)g.a.Mo * g.a.No * C * B * sizeof(float));
  if (bytes) *bytes = n;
  return ADMMTV_OK;
}

int admmtv_ssim_window_forward(int M, int N, int C, int B, int device, const float* x, const float* y, const float* u,
                               const float* v, int L1, int L2, int R, float peakval, int as_loss, float* out, void* workspace,
                               int with_grad, void* stream) {
  if (!u || !v) return ADMMTV_ERR_NULL;
  SsimGenArgs g{};
  int rc = ssim_gen_args(g, M, N, C, B, u, v, L1, L2, R);
  if (rc) return rc;
  if (!x || !y || !out || !workspace) return ADMMTV_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return ADMMTV_ERR_ALIGN;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SsimArgs& a = g.a;
  a.x = x; a.y = y;
  a.C1 = (peakval * 0.01f) * (peakval * 0.01f);  // ssim.jl:101-102
  a.C2 = (peakval * 0.03f) * (peakval * 0.03f);
  a.tiles_i = (a.Mo + SS_T - 1) / SS_T; a.tiles_j = (a.No + SS_T - 1) / SS_T;
  a.acc = reinterpret_cast<double*>(workspace);
  a.maps = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + up256(sizeof(double)));
  a.with_grad = with_grad ? 1 : 0;
  a.as_loss = as_loss ? 1 : 0;
  a.out = out;
  cudaError_t e = cudaMemsetAsync(a.acc, 0, sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  const size_t nblk = (size_t)a.tiles_i * a.tiles_j * C * B;
  if (nblk > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  ADMMTV_LAUNCH(k_ssim_fwd_gen, dim3((unsigned)nblk), dim3(SS_NT), 0, st, g);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  ADMMTV_LAUNCH(k_ssim_finalize, dim3(1), dim3(1), 0, st, a);
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  return ADMMTV_OK;
}

int admmtv_ssim_window_backward(int M, int N, int C, int B, int device, const float* x, const float* y, const float* u,
                                const float* v, int L1, int L2, int R, int as_loss, const float* outbar, const void* workspace,
                                float* xbar, void* stream) {
  if (!u || !v) return ADMMTV_ERR_NULL;
  SsimGenArgs g{};
  int rc = ssim_gen_args(g, M, N, C, B, u, v, L1, L2, R);
  if (rc) return rc;
  if (!x || !y || !outbar || !workspace || !xbar) return ADMMTV_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return ADMMTV_ERR_ALIGN;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SsimArgs& a = g.a;
  a.x = x; a.y = y;
  a.tiles_i = (M + SS_T - 1) / SS_T; a.tiles_j = (N + SS_T - 1) / SS_T;
  unsigned char* ws = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(workspace));
  a.acc = reinterpret_cast<double*>(ws);
  a.maps = reinterpret_cast<float*>(ws + up256(sizeof(double)));
  a.with_grad = 1;
  a.as_loss = as_loss ? 1 : 0;
  a.outbar = outbar;
  a.out = xbar;
  const size_t nblk = (size_t)a.tiles_i * a.tiles_j * C * B;
  if (nblk > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  ADMMTV_LAUNCH(k_ssim_bwd_gen, dim3((unsigned)nblk), dim3(SS_NT), 0, st, g);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADMMTV_OK : (int)e;
}

int admmtv_pad_symmetric(int M, int N, int planes, int lo1, int hi1, int lo2, int hi2, int device, const float* src,
                         float* dst, void* stream) {
  PadArgs a{};
  int rc = pad_args(a, M, N, planes, lo1, hi1, lo2, hi2);
  if (rc) return rc;
  if (!src || !dst) return ADMMTV_ERR_NULL;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  a.src = src; a.dst = dst;
  const size_t total = (size_t)(M + lo1 + hi1) * (N + lo2 + hi2) * planes;
  ADMMTV_LAUNCH(k_pad_symmetric, dim3(pad_grid(total)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADMMTV_OK : (int)e;
}

int admmtv_pad_symmetric_adjoint(int M, int N, int planes, int lo1, int hi1, int lo2, int hi2, int device,
                                 const float* padded_bar, float* src_bar, void* stream) {
  PadArgs a{};
  int rc = pad_args(a, M, N, planes, lo1, hi1, lo2, hi2);
  if (rc) return rc;
  if (!padded_bar || !src_bar) return ADMMTV_ERR_NULL;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  a.src = padded_bar; a.dst = src_bar;
  const size_t total = (size_t)M * N * planes;
  ADMMTV_LAUNCH(k_pad_symmetric_adj, dim3(pad_grid(total)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADMMTV_OK : (int)e;
}

int admmtv_batch_from_n0f8(int M, int N, int C, int B, int device, const uint8_t* src, int64_t stride_c, int64_t stride_i,
                           int64_t stride_j, int64_t stride_b, float* dst, void* stream) {
  int rc = check_shape(M, N, C, B);
  if (rc) return rc;
  if (!src || !dst) return ADMMTV_ERR_NULL;
  if (stride_c < 0 || stride_i < 0 || stride_j < 0 || stride_b < 0) return ADMMTV_ERR_SHAPE;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  {
    // contiguous layouts take the vectorised kernels (bit-identical results)
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t plane = (size_t)M * N, n = plane * C * B;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    if (aligned && stride_i == 1 && stride_j == M && stride_c == (int64_t)plane && stride_b == (int64_t)(plane * C)) {
      const size_t n16 = n / 16, nb = (n16 + 255) / 256;
      ADMMTV_LAUNCH(k_batch_flat_n0f8, dim3((unsigned)(nb < 148 * 32 ? (nb ? nb : 1) : 148 * 32)), dim3(256), 0, st, src, dst, n16, n);
      cudaError_t e = cudaGetLastError();
      return e == cudaSuccess ? ADMMTV_OK : (int)e;
    }
    if (aligned && stride_c == 1 && stride_i == C && stride_j == (int64_t)C * M && stride_b % 4 == 0 && plane % 4 == 0 &&
        (C == 1 || C == 3 || C == 4)) {
      const size_t px4 = plane / 4, tot = px4 * B, nb = (tot + 255) / 256;
      const dim3 grid((unsigned)(nb < 148 * 32 ? (nb ? nb : 1) : 148 * 32));
      if (C == 1) ADMMTV_LAUNCH(k_batch_interleaved_n0f8<1>, grid, dim3(256), 0, st, src, dst, px4, plane, (long long)stride_b, B);
      else if (C == 3) ADMMTV_LAUNCH(k_batch_interleaved_n0f8<3>, grid, dim3(256), 0, st, src, dst, px4, plane, (long long)stride_b, B);
      else ADMMTV_LAUNCH(k_batch_interleaved_n0f8<4>, grid, dim3(256), 0, st, src, dst, px4, plane, (long long)stride_b, B);
      cudaError_t e = cudaGetLastError();
      return e == cudaSuccess ? ADMMTV_OK : (int)e;
    }
  }
  BatchArgs a{};
  a.src = src; a.dst = dst; a.M = M; a.N = N; a.C = C; a.B = B;
  a.tiles_i = (M + BA_T - 1) / BA_T; a.tiles_j = (N + BA_T - 1) / BA_T;
  a.sc = stride_c; a.si = stride_i; a.sj = stride_j; a.sb = stride_b;
  const size_t nblk = (size_t)a.tiles_i * a.tiles_j * B;
  if (nblk > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  ADMMTV_LAUNCH(k_batch_from_n0f8, dim3((unsigned)nblk), dim3(BA_NT), 0, reinterpret_cast<cudaStream_t>(stream), a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADMMTV_OK : (int)e;
}

int admmtv_batch_gather_n0f8(int M, int N, int C, int B, int device, const uint8_t* base, const int64_t* offsets,
                             int64_t stride_c, int64_t stride_i, int64_t stride_j, float* dst, void* stream) {
  int rc = check_shape(M, N, C, B);
  if (rc) return rc;
  if (!base || !offsets || !dst) return ADMMTV_ERR_NULL;
  if (stride_c < 0 || stride_i < 0 || stride_j < 0) return ADMMTV_ERR_SHAPE;
  DevGuard guard(device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  BatchArgs a{};
  a.src = base; a.dst = dst; a.M = M; a.N = N; a.C = C; a.B = B;
  a.tiles_i = (M + BA_T - 1) / BA_T; a.tiles_j = (N + BA_T - 1) / BA_T;
  a.sc = stride_c; a.si = stride_i; a.sj = stride_j; a.sb = 0;
  a.offsets = reinterpret_cast<const long long*>(offsets);
  const size_t nblk = (size_t)a.tiles_i * a.tiles_j * B;
  if (nblk > 0x7fffffffu) return ADMMTV_ERR_SHAPE;
  ADMMTV_LAUNCH(k_batch_from_n0f8, dim3((unsigned)nblk), dim3(BA_NT), 0, reinterpret_cast<cudaStream_t>(stream), a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADMMTV_OK : (int)e;
}

}  // extern "C"
