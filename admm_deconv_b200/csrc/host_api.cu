// host_api.cu -- the host-buffer entry points of include/admmtv_host.h: a thin, stream-pipelined layer over the
// device-pointer C ABI (admmtv_forward / admmtv_backward).  Everything still runs on the GPU.
#include "../../include/admmtv_host.h"
#include "../../include/admmtv_batch.h"

#include <cuda_runtime.h>
#include <new>
#include <stdint.h>
#include <string.h>

namespace {

inline size_t al256(size_t n) { return (n + 255) & ~size_t(255); }

struct Slot {
  float *y, *target, *x;       // device
  cudaEvent_t in_ready, compute_done, out_done;
  float* loss_out;             // host destination of the pending step's loss (or null)
  bool pending;
};

}  // namespace

struct admmtv_host_session {
  admmtv_desc d;
  int training, G, nh, ngrad, PS;
  size_t nimg;                 // floats per (M,N,P,B) array
  size_t out_img;              // floats of x_out (== nimg unless channel-concat/shared input)
  size_t in_img;               // floats of y
  unsigned char* arena;
  bool own_arena, own_compute;
  cudaStream_t copy_in, compute, copy_out;
  Slot slot[2];
  float *h[2], *lambda[2], *rho[2], *bias[2];   // device parameters, one set per slot (uploaded ahead of the slot's images)
  float *ybar, *packed;             // training
  double* loss_acc;                 // device
  double* loss_host[2];             // pinned
  void *ws_fwd, *ws_bwd, *ckpt;
  int prev_device;
};

namespace {

struct Layout {
  size_t y[2], t[2], x[2], h[2], lam[2], rho[2], bias[2], ybar, packed, loss, ws_fwd, ws_bwd, ckpt, total;
};

inline int param_entries(const admmtv_desc* d) { return (d->flags & ADMMTV_FLAG_PER_ITER_PARAMS) ? d->iters : 1; }

int plan(const admmtv_desc* d, int training, Layout& L, size_t& in_img, size_t& out_img, int& G, int& nh, int& ngrad) {
  int rc = admmtv_check(d);
  if (rc) return rc;
  G = d->groups > 1 ? d->groups : 1;
  nh = d->kh * d->kw;
  const int PS = param_entries(d);
  ngrad = nh * G + 2 * G * PS + (d->has_bias ? G : 0);
  const size_t plane = (size_t)d->M * d->N;
  out_img = plane * d->P * d->B;
  in_img = (d->flags & ADMMTV_FLAG_SHARED_INPUT) ? plane * d->P * (d->B / G) : out_img;
  size_t fwd = 0, ck = 0, bwd = 0;
  if ((rc = admmtv_workspace_bytes(d, &fwd, &ck, &bwd))) return rc;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o += al256(bytes); return r; };
  for (int s = 0; s < 2; ++s) {
    L.y[s] = take(in_img * 4);
    L.t[s] = take(training ? out_img * 4 : 0);
    L.x[s] = take(out_img * 4);
  }
  for (int s = 0; s < 2; ++s) {
    L.h[s] = take((size_t)(nh > 0 ? nh : 1) * G * 4);
    L.lam[s] = take((size_t)G * PS * 4);
    L.rho[s] = take((size_t)G * PS * 4);
    L.bias[s] = take((size_t)G * 4);
  }
  L.ybar = take(training ? in_img * 4 : 0);
  L.packed = take((size_t)ngrad * 4);
  L.loss = take(16);
  L.ws_fwd = take(fwd);
  L.ws_bwd = take(training ? bwd : 0);
  L.ckpt = take(training ? ck : 0);
  L.total = o;
  return 0;
}

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

#define HCHECK(expr)                          \
  do {                                        \
    cudaError_t e__ = (expr);                 \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

// Parameters host -> device (tiny).  They travel on the COPY-IN stream in front of the slot's images and into the slot's own
// parameter set: on the compute stream they would queue on the host-to-device copy engine behind the NEXT step's images,
// which are issued while this step is still waiting to start.
int upload_params(admmtv_host_session* s, int slot, const float* h, const float* lambda, const float* rho, const float* bias) {
  if (s->nh > 0) HCHECK(cudaMemcpyAsync(s->h[slot], h, (size_t)s->nh * s->G * 4, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaMemcpyAsync(s->lambda[slot], lambda, (size_t)s->G * s->PS * 4, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaMemcpyAsync(s->rho[slot], rho, (size_t)s->G * s->PS * 4, cudaMemcpyHostToDevice, s->copy_in));
  if (s->d.has_bias) HCHECK(cudaMemcpyAsync(s->bias[slot], bias, (size_t)s->G * 4, cudaMemcpyHostToDevice, s->copy_in));
  return 0;
}
// the persisted clamp (deconv_admm.jl:216-219) back to the caller's arrays.  These few bytes travel on the COMPUTE stream
// (like the gradients and the loss): stream order keeps them behind the kernels that wrote them; only the image-sized
// transfers use the copy streams.
int download_params(admmtv_host_session* s, int slot, float* h, float* lambda, float* rho) {
  if (s->d.flags & ADMMTV_FLAG_NO_CLAMP) return 0;
  if (s->nh > 0) HCHECK(cudaMemcpyAsync(h, s->h[slot], (size_t)s->nh * s->G * 4, cudaMemcpyDeviceToHost, s->compute));
  HCHECK(cudaMemcpyAsync(lambda, s->lambda[slot], (size_t)s->G * s->PS * 4, cudaMemcpyDeviceToHost, s->compute));
  HCHECK(cudaMemcpyAsync(rho, s->rho[slot], (size_t)s->G * s->PS * 4, cudaMemcpyDeviceToHost, s->compute));
  return 0;
}

}  // namespace

extern "C" {

int admmtv_host_grad_floats(const admmtv_desc* d) {
  if (admmtv_check(d)) return 0;
  const int G = d->groups > 1 ? d->groups : 1;
  return d->kh * d->kw * G + 2 * G * param_entries(d) + (d->has_bias ? G : 0);
}

int admmtv_host_session_bytes(const admmtv_desc* d, int training, size_t* device_bytes) {
  if (!device_bytes) return ADMMTV_ERR_NULL;
  Layout L;
  size_t a, b;
  int G, nh, ng;
  int rc = plan(d, training, L, a, b, G, nh, ng);
  if (rc) return rc;
  *device_bytes = L.total;
  return ADMMTV_OK;
}

int admmtv_host_session_create(const admmtv_desc* d, int training, void* device_arena, void* compute_stream,
                               admmtv_host_session** out) {
  if (!out) return ADMMTV_ERR_NULL;
  *out = nullptr;
  Layout L;
  size_t in_img, out_img;
  int G, nh, ng;
  int rc = plan(d, training, L, in_img, out_img, G, nh, ng);
  if (rc) return rc;
  if (reinterpret_cast<uintptr_t>(device_arena) & 255) return ADMMTV_ERR_ALIGN;
  DevGuard guard(d->device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  admmtv_host_session* s = new (std::nothrow) admmtv_host_session();
  if (!s) return (int)cudaErrorMemoryAllocation;
  memset(s, 0, sizeof(*s));
  s->d = *d;
  s->training = training ? 1 : 0;
  s->G = G; s->nh = nh; s->ngrad = ng; s->in_img = in_img; s->out_img = out_img; s->PS = param_entries(d);
  s->own_arena = device_arena == nullptr;
  s->arena = reinterpret_cast<unsigned char*>(device_arena);
  cudaError_t e = cudaSuccess;
  if (s->own_arena) e = cudaMalloc((void**)&s->arena, L.total);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_out, cudaStreamNonBlocking);
  s->own_compute = compute_stream == nullptr;
  s->compute = reinterpret_cast<cudaStream_t>(compute_stream);
  if (e == cudaSuccess && s->own_compute) e = cudaStreamCreateWithFlags(&s->compute, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    Slot& sl = s->slot[i];
    sl.y = reinterpret_cast<float*>(s->arena + L.y[i]);
    sl.target = reinterpret_cast<float*>(s->arena + L.t[i]);
    sl.x = reinterpret_cast<float*>(s->arena + L.x[i]);
    e = cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->loss_host[i], sizeof(double));
  }
  if (e != cudaSuccess) {
    admmtv_host_session_destroy(s);
    return (int)e;
  }
  for (int i = 0; i < 2; ++i) {
    s->h[i] = reinterpret_cast<float*>(s->arena + L.h[i]);
    s->lambda[i] = reinterpret_cast<float*>(s->arena + L.lam[i]);
    s->rho[i] = reinterpret_cast<float*>(s->arena + L.rho[i]);
    s->bias[i] = reinterpret_cast<float*>(s->arena + L.bias[i]);
  }
  s->ybar = reinterpret_cast<float*>(s->arena + L.ybar);
  s->packed = reinterpret_cast<float*>(s->arena + L.packed);
  s->loss_acc = reinterpret_cast<double*>(s->arena + L.loss);
  s->ws_fwd = s->arena + L.ws_fwd;
  s->ws_bwd = s->arena + L.ws_bwd;
  s->ckpt = s->arena + L.ckpt;
  *out = s;
  return ADMMTV_OK;
}

int admmtv_host_session_destroy(admmtv_host_session* s) {
  if (!s) return ADMMTV_OK;
  DevGuard guard(s->d.device);
  if (s->copy_in) cudaStreamSynchronize(s->copy_in);
  if (s->compute) cudaStreamSynchronize(s->compute);
  if (s->copy_out) cudaStreamSynchronize(s->copy_out);
  for (int i = 0; i < 2; ++i) {
    if (s->slot[i].in_ready) cudaEventDestroy(s->slot[i].in_ready);
    if (s->slot[i].compute_done) cudaEventDestroy(s->slot[i].compute_done);
    if (s->slot[i].out_done) cudaEventDestroy(s->slot[i].out_done);
    if (s->loss_host[i]) cudaFreeHost(s->loss_host[i]);
  }
  if (s->copy_in) cudaStreamDestroy(s->copy_in);
  if (s->copy_out) cudaStreamDestroy(s->copy_out);
  if (s->own_compute && s->compute) cudaStreamDestroy(s->compute);
  if (s->own_arena && s->arena) cudaFree(s->arena);
  delete s;
  return ADMMTV_OK;
}

int admmtv_host_pin(void* p, size_t bytes) {
  if (!p) return ADMMTV_ERR_NULL;
  return (int)cudaHostRegister(p, bytes, cudaHostRegisterPortable);
}
int admmtv_host_unpin(void* p) {
  if (!p) return ADMMTV_ERR_NULL;
  return (int)cudaHostUnregister(p);
}

int admmtv_host_launches(const admmtv_host_session* s, int training) {
  if (!s) return 0;
  int n = admmtv_forward_launches(&s->d, training);
  if (training) n += admmtv_backward_launches(&s->d);
  return n;
}

int admmtv_host_forward_enqueue(admmtv_host_session* s, int slot, const float* y, float* h, float* lambda, float* rho,
                                const float* bias, float* x_out) {
  if (!s || !y || !lambda || !rho || !x_out) return ADMMTV_ERR_NULL;
  if (s->nh > 0 && !h) return ADMMTV_ERR_NULL;   // checked before any copy
  if (s->d.has_bias && !bias) return ADMMTV_ERR_NULL;
  if (slot < 0 || slot > 1) return ADMMTV_ERR_ENUM;
  DevGuard guard(s->d.device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  Slot& sl = s->slot[slot];
  // the slot's previous use was waited for by the caller, so its device buffers (images and parameter set) are free
  int rc = upload_params(s, slot, h, lambda, rho, bias);
  if (rc) return rc;
  HCHECK(cudaMemcpyAsync(sl.y, y, s->in_img * 4, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaEventRecord(sl.in_ready, s->copy_in));
  HCHECK(cudaStreamWaitEvent(s->compute, sl.in_ready, 0));
  rc = admmtv_forward(&s->d, sl.y, s->nh > 0 ? s->h[slot] : nullptr, s->lambda[slot], s->rho[slot],
                      s->d.has_bias ? s->bias[slot] : nullptr, sl.x, s->ws_fwd, nullptr, s->compute);
  if (rc) return rc;
  if ((rc = download_params(s, slot, h, lambda, rho))) return rc;
  HCHECK(cudaEventRecord(sl.compute_done, s->compute));
  HCHECK(cudaStreamWaitEvent(s->copy_out, sl.compute_done, 0));
  HCHECK(cudaMemcpyAsync(x_out, sl.x, s->out_img * 4, cudaMemcpyDeviceToHost, s->copy_out));
  HCHECK(cudaEventRecord(sl.out_done, s->copy_out));
  sl.loss_out = nullptr;
  sl.pending = true;
  return ADMMTV_OK;
}

int admmtv_host_forward_enqueue_n0f8(admmtv_host_session* s, int slot, const uint8_t* y, int64_t stride_c, int64_t stride_i,
                                     int64_t stride_j, int64_t stride_b, float* h, float* lambda, float* rho, const float* bias,
                                     float* x_out) {
  if (!s || !y || !lambda || !rho || !x_out) return ADMMTV_ERR_NULL;
  if (s->nh > 0 && !h) return ADMMTV_ERR_NULL;
  if (s->d.has_bias && !bias) return ADMMTV_ERR_NULL;
  if (slot < 0 || slot > 1) return ADMMTV_ERR_ENUM;
  if (stride_c < 0 || stride_i < 0 || stride_j < 0 || stride_b < 0) return ADMMTV_ERR_SHAPE;
  DevGuard guard(s->d.device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  Slot& sl = s->slot[slot];
  // the slot's output buffer x (written by the forward only after the conversion kernel has run) stages the bytes
  const int Bin = (int)(s->in_img / ((size_t)s->d.M * s->d.N * s->d.P));
  const size_t ny = (size_t)((int64_t)(s->d.P - 1) * stride_c + (int64_t)(s->d.M - 1) * stride_i + (int64_t)(s->d.N - 1) * stride_j +
                             (int64_t)(Bin - 1) * stride_b + 1);
  if (ny > s->out_img * 4) return ADMMTV_ERR_SHAPE;
  uint8_t* sy = reinterpret_cast<uint8_t*>(sl.x);
  int rc = upload_params(s, slot, h, lambda, rho, bias);
  if (rc) return rc;
  HCHECK(cudaMemcpyAsync(sy, y, ny, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaEventRecord(sl.in_ready, s->copy_in));
  HCHECK(cudaStreamWaitEvent(s->compute, sl.in_ready, 0));
  if ((rc = admmtv_batch_from_n0f8(s->d.M, s->d.N, s->d.P, Bin, s->d.device, sy, stride_c, stride_i, stride_j, stride_b, sl.y, s->compute)))
    return rc;
  rc = admmtv_forward(&s->d, sl.y, s->nh > 0 ? s->h[slot] : nullptr, s->lambda[slot], s->rho[slot],
                      s->d.has_bias ? s->bias[slot] : nullptr, sl.x, s->ws_fwd, nullptr, s->compute);
  if (rc) return rc;
  if ((rc = download_params(s, slot, h, lambda, rho))) return rc;
  HCHECK(cudaEventRecord(sl.compute_done, s->compute));
  HCHECK(cudaStreamWaitEvent(s->copy_out, sl.compute_done, 0));
  HCHECK(cudaMemcpyAsync(x_out, sl.x, s->out_img * 4, cudaMemcpyDeviceToHost, s->copy_out));
  HCHECK(cudaEventRecord(sl.out_done, s->copy_out));
  sl.loss_out = nullptr;
  sl.pending = true;
  return ADMMTV_OK;
}

int admmtv_host_train_step_enqueue(admmtv_host_session* s, int slot, const float* y, const float* target, float* h,
                                   float* lambda, float* rho, const float* bias, float* grads_out, float* loss_out,
                                   float* ybar_out, const admmtv_hooks* hooks) {
  if (!s || !y || !target || !lambda || !rho || !grads_out) return ADMMTV_ERR_NULL;
  if (!s->training) return ADMMTV_ERR_ENUM;
  if (s->nh > 0 && !h) return ADMMTV_ERR_NULL;
  if (s->d.has_bias && !bias) return ADMMTV_ERR_NULL;
  if (slot < 0 || slot > 1) return ADMMTV_ERR_ENUM;
  DevGuard guard(s->d.device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  Slot& sl = s->slot[slot];
  int rc = upload_params(s, slot, h, lambda, rho, bias);
  if (rc) return rc;
  HCHECK(cudaMemcpyAsync(sl.y, y, s->in_img * 4, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaMemcpyAsync(sl.target, target, s->out_img * 4, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaEventRecord(sl.in_ready, s->copy_in));
  HCHECK(cudaStreamWaitEvent(s->compute, sl.in_ready, 0));
  if ((rc = admmtv_mse_train_step(&s->d, sl.y, sl.target, s->nh > 0 ? s->h[slot] : nullptr, s->lambda[slot], s->rho[slot],
                                  s->d.has_bias ? s->bias[slot] : nullptr, sl.x, s->ybar, s->packed, s->loss_acc, s->ws_fwd,
                                  s->ckpt, s->ws_bwd, s->compute, hooks)))
    return rc;
  HCHECK(cudaMemcpyAsync(grads_out, s->packed, (size_t)s->ngrad * 4, cudaMemcpyDeviceToHost, s->compute));
  HCHECK(cudaMemcpyAsync(s->loss_host[slot], s->loss_acc, sizeof(double), cudaMemcpyDeviceToHost, s->compute));
  if ((rc = download_params(s, slot, h, lambda, rho))) return rc;
  HCHECK(cudaEventRecord(sl.compute_done, s->compute));
  HCHECK(cudaStreamWaitEvent(s->copy_out, sl.compute_done, 0));
  if (ybar_out) {
    HCHECK(cudaMemcpyAsync(ybar_out, s->ybar, s->in_img * 4, cudaMemcpyDeviceToHost, s->copy_out));
    HCHECK(cudaEventRecord(sl.out_done, s->copy_out));
    HCHECK(cudaStreamWaitEvent(s->compute, sl.out_done, 0));   // the next step's backward overwrites ybar
  } else {
    HCHECK(cudaEventRecord(sl.out_done, s->copy_out));
  }
  sl.loss_out = loss_out;
  sl.pending = true;
  return ADMMTV_OK;
}

int admmtv_host_train_step_enqueue_n0f8(admmtv_host_session* s, int slot, const uint8_t* y, const uint8_t* target,
                                        int64_t stride_c, int64_t stride_i, int64_t stride_j, int64_t stride_b, float* h,
                                        float* lambda, float* rho, const float* bias, float* grads_out, float* loss_out,
                                        const admmtv_hooks* hooks) {
  if (!s || !y || !target || !lambda || !rho || !grads_out) return ADMMTV_ERR_NULL;
  if (!s->training) return ADMMTV_ERR_ENUM;
  if (s->nh > 0 && !h) return ADMMTV_ERR_NULL;
  if (s->d.has_bias && !bias) return ADMMTV_ERR_NULL;
  if (slot < 0 || slot > 1) return ADMMTV_ERR_ENUM;
  if (stride_c < 0 || stride_i < 0 || stride_j < 0 || stride_b < 0) return ADMMTV_ERR_SHAPE;
  DevGuard guard(s->d.device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  Slot& sl = s->slot[slot];
  // The slot's output buffer x (4 bytes per sample, written by the forward only after the conversion kernels have run) stages
  // the two byte arrays: in_img + out_img <= 4 out_img bytes.  The byte extent of a strided source is its largest offset + 1.
  const int Bin = (int)(s->in_img / ((size_t)s->d.M * s->d.N * s->d.P));
  auto extent = [&](int B) {
    return (size_t)((int64_t)(s->d.P - 1) * stride_c + (int64_t)(s->d.M - 1) * stride_i + (int64_t)(s->d.N - 1) * stride_j +
                    (int64_t)(B - 1) * stride_b + 1);
  };
  const size_t ny = extent(Bin), nt = extent(s->d.B);
  if (al256(ny) + nt > s->out_img * 4) return ADMMTV_ERR_SHAPE;   // padded strides larger than 4 bytes per sample
  uint8_t* sy = reinterpret_cast<uint8_t*>(sl.x);
  uint8_t* st = sy + al256(ny);
  int rc = upload_params(s, slot, h, lambda, rho, bias);
  if (rc) return rc;
  HCHECK(cudaMemcpyAsync(sy, y, ny, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaMemcpyAsync(st, target, nt, cudaMemcpyHostToDevice, s->copy_in));
  HCHECK(cudaEventRecord(sl.in_ready, s->copy_in));
  HCHECK(cudaStreamWaitEvent(s->compute, sl.in_ready, 0));
  if ((rc = admmtv_batch_from_n0f8(s->d.M, s->d.N, s->d.P, Bin, s->d.device, sy, stride_c, stride_i, stride_j, stride_b, sl.y, s->compute)))
    return rc;
  if ((rc = admmtv_batch_from_n0f8(s->d.M, s->d.N, s->d.P, s->d.B, s->d.device, st, stride_c, stride_i, stride_j, stride_b, sl.target,
                                   s->compute)))
    return rc;
  if ((rc = admmtv_mse_train_step(&s->d, sl.y, sl.target, s->nh > 0 ? s->h[slot] : nullptr, s->lambda[slot], s->rho[slot],
                                  s->d.has_bias ? s->bias[slot] : nullptr, sl.x, s->ybar, s->packed, s->loss_acc, s->ws_fwd,
                                  s->ckpt, s->ws_bwd, s->compute, hooks)))
    return rc;
  HCHECK(cudaMemcpyAsync(grads_out, s->packed, (size_t)s->ngrad * 4, cudaMemcpyDeviceToHost, s->compute));
  HCHECK(cudaMemcpyAsync(s->loss_host[slot], s->loss_acc, sizeof(double), cudaMemcpyDeviceToHost, s->compute));
  if ((rc = download_params(s, slot, h, lambda, rho))) return rc;
  HCHECK(cudaEventRecord(sl.compute_done, s->compute));
  HCHECK(cudaStreamWaitEvent(s->copy_out, sl.compute_done, 0));
  HCHECK(cudaEventRecord(sl.out_done, s->copy_out));
  sl.loss_out = loss_out;
  sl.pending = true;
  return ADMMTV_OK;
}

int admmtv_mse_train_step(const admmtv_desc* d, const float* y, const float* target, float* h, float* lambda, float* rho,
                          const float* bias, float* x_out, float* ybar, float* grads_packed, double* loss_sum,
                          void* ws_fwd, void* ckpt, void* ws_bwd, void* stream, const admmtv_hooks* hooks) {
  int rc = admmtv_check(d);
  if (rc) return rc;
  if (!y || !target || !lambda || !rho || !x_out || !ybar || !grads_packed || !loss_sum || !ws_fwd || !ckpt || !ws_bwd)
    return ADMMTV_ERR_NULL;
  DevGuard guard(d->device);
  if (!guard.ok) return ADMMTV_ERR_NO_DEVICE;
  const int G = d->groups > 1 ? d->groups : 1, nh = d->kh * d->kw, PS = param_entries(d);
  if ((rc = admmtv_forward(d, y, h, lambda, rho, bias, x_out, ws_fwd, ckpt, stream))) return rc;
  // loss = mean((x - target)^2): its pullback seed 2 (x - target) / numel is formed inside the backward's first kernel
  float* hbar = grads_packed;
  float* lbar = grads_packed + (size_t)nh * G;
  float* rbar = lbar + (size_t)G * PS;
  float* bbar = d->has_bias ? rbar + (size_t)G * PS : nullptr;
  if ((rc = admmtv_backward_mse(d, target, x_out, y, h, lambda, rho, ckpt, ybar, nh > 0 ? hbar : nullptr, lbar, rbar, bbar, loss_sum,
                                ws_bwd, stream)))
    return rc;
  if (hooks && hooks->allreduce_sum) {   // the data-parallel gradient all-reduce, stream-ordered
    const size_t ng = (size_t)nh * G + 2 * (size_t)G * PS + (d->has_bias ? G : 0);
    if ((rc = hooks->allreduce_sum(grads_packed, ng, stream, hooks->user))) return rc;
  }
  return ADMMTV_OK;
}

int admmtv_host_wait(admmtv_host_session* s, int slot) {
  if (!s) return ADMMTV_ERR_NULL;
  if (slot < 0 || slot > 1) return ADMMTV_ERR_ENUM;
  Slot& sl = s->slot[slot];
  if (!sl.pending) return ADMMTV_OK;
  DevGuard guard(s->d.device);
  HCHECK(cudaEventSynchronize(sl.out_done));
  if (sl.loss_out) *sl.loss_out = (float)(*s->loss_host[slot] / (double)s->out_img);
  sl.pending = false;
  return ADMMTV_OK;
}

}  // extern "C"
