/*
 * admmtv.h -- C ABI of the B200-native ADMM-TV deconvolution layer.
 *
 * This is the drop-in boundary for ONE hot path of georgegrosu1/admm-deconv:
 *
 *   reference interface replaced                         entry point here
 *   ---------------------------------------------------  ---------------------------------
 *   tvd_fft(y, λ, ρ, h, isotropic, maxit)                admmtv_forward  (activation=identity,
 *     src/ops/ops.jl:181-188 -> tvd_fft_gpu :99-178        has_bias=0, ADMMTV_FLAG_NO_CLAMP)
 *   (d::Admm)(x)  clamp -> tvd_fft -> +bias -> σ          admmtv_forward
 *     src/layers/deconv_admm.jl:215-225
 *   Zygote pullback of the two above (no reference code;  admmtv_backward
 *     tape through ops.jl:166-174)
 *   CUDA.zeros / temporaries of ops.jl:128-141            admmtv_workspace_bytes (caller owns memory)
 *   tvd_fft on a CPU Array (ops.jl:187 -> tvd_fft_cpu)    admmtv_forward_host  (host buffers, copies inside)
 *
 * Conventions
 *   - Arrays are Julia arrays: (M,N,P,B) column-major, element (i,j,p,b) at i + M*(j + N*(p + P*b)).
 *     h is (kh,kw) column-major (the (kh,kw,1,1) Flux.convfilter weight).  All fp32.
 *   - Unless the name ends in _host, every data pointer is a DEVICE pointer on `desc->device`
 *     (Julia passes CuPtr{Cfloat}); `stream` is a cudaStream_t passed as void*.
 *   - Calls only enqueue work on `stream` and return; they never synchronise.
 *   - Return value: 0 ok; <0 invalid argument / unsupported size (nothing was launched);
 *     >0 a cudaError_t.  admmtv_strerror() explains either.  Nothing throws across the boundary.
 *   - The library holds no mutable global state; every table lives in the caller's workspace.
 *   - There is no CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef ADMMTV_H
#define ADMMTV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMMTV_VERSION 100 /* 0.1.0 */

/* activation enum: the σ of deconv_admm.jl:224 as used by net_build.jl (:8,107-119,155) */
enum { ADMMTV_ACT_IDENTITY = 0, ADMMTV_ACT_RELU = 1, ADMMTV_ACT_RELU6 = 2, ADMMTV_ACT_RELU1 = 3 };

/* flags */
enum {
  ADMMTV_FLAG_NO_CLAMP = 1,      /* skip deconv_admm.jl:216-219 (bare tvd_fft call) */
  ADMMTV_FLAG_NOGRAD_REPEAT = 2, /* train.jl:10 variant: no ∂weight through the spatial H^T y path */
  /* grouped calls (desc.groups > 1), see below */
  ADMMTV_FLAG_SHARED_INPUT = 4,  /* every group reads the same y (M,N,P,B/groups) */
  ADMMTV_FLAG_CHANNEL_CONCAT = 8, /* x_out is (M,N,groups*P,B/groups): group g in channels [gP,(g+1)P) */
  /* accepted and ignored (kept for ABI stability): the isotropic per-pixel terms are always reduced by a small
   * kernel per iteration that adds the plane pairs' shares in a fixed order, so the isotropic path is
   * bit-reproducible run to run (no floating-point atomics on it) */
  ADMMTV_FLAG_ISO_PRECOMPUTE = 16,
  ADMMTV_FLAG_ISO_INLINE = 32,
  /* EXTENSION (BASELINE configs[2] "learned-rho/lambda iterations"): one (lambda_k, rho_k) per unrolled
   * iteration instead of the reference's single pair (ops.jl:20: tau is one scalar for all iterations).  lambda and rho
   * then hold `iters` floats per group, entry k-1 for iteration k: x_k = F^-1(C(rho_k) F(H^T y + rho_k D^T(z_{k-1} - u_{k-1}))),
   * z_k = shrink(D x_k + u_{k-1}, lambda_k / rho_k).  admmtv_backward returns `iters` lambdabar / rhobar per group (the last
   * lambdabar is 0: the final z-update is dead, ops.jl:89-91).  With every entry equal the result is bit-identical to the
   * shared-parameter call. */
  ADMMTV_FLAG_PER_ITER_PARAMS = 64,
  /* tuning / testing: planes of 32^2, 64^2, 128^2 normally run every iteration of an anisotropic inference call inside one
   * persistent kernel per plane pair (kernels_small.cuh); this flag keeps the general two-launch-per-iteration path */
  ADMMTV_FLAG_NO_SMALL = 128
};

/* error codes (<0) */
enum {
  ADMMTV_OK = 0,
  ADMMTV_ERR_NULL = -1,
  ADMMTV_ERR_SHAPE = -2,        /* non-positive dims, kernel larger than image, ... */
  ADMMTV_ERR_UNSUPPORTED = -3,  /* M or N above 4096.  Every size 1..4096 is accepted (the reference's FFTW path takes
                                 * any size, ops.jl:26,86): when both M and N are planned FFT lengths -- 32, 64, ..., 4096
                                 * or 96, 160, 192, 320, 384, 480, 640, 768, 960, 1280, 1536, 1920 -- the tuned fused
                                 * kernels run; any other size takes the generic kernels (same results, slower) */
  ADMMTV_ERR_ITERS = -4,
  ADMMTV_ERR_ENUM = -5,
  ADMMTV_ERR_ALIGN = -6,        /* workspace / checkpoint not 256-byte aligned */
  ADMMTV_ERR_NO_DEVICE = -7
};

typedef struct admmtv_desc {
  int32_t M, N, P, B;      /* size(y)                                   ops.jl:100 */
  int32_t kh, kw;          /* size(h)[1:2]; 0,0 = empty h (H = identity) ops.jl:104 */
  int32_t iters;           /* maxit / d.iters                            ops.jl:166 */
  int32_t iso;             /* isotropic: 0 = ST (ops.jl:9), 1 = BT (ops.jl:10) */
  int32_t activation;      /* ADMMTV_ACT_*                               deconv_admm.jl:224 */
  int32_t has_bias;        /* d.bias is a 1-vector (1) or `false` (0)    deconv_admm.jl:222 */
  int32_t device;          /* CUDA device ordinal */
  int32_t flags;           /* ADMMTV_FLAG_* */
  float creg;              /* clamp floor of λ, ρ                        deconv_admm.jl:216-217 */
  int32_t groups;          /* 0/1: one call (the reference).  G > 1: G independent calls of identical shape
                            * batched into one launch sequence -- an EXTENSION: image b of the
                            * B images belongs to group b / (B/G); λ, ρ, bias hold G floats, h holds G kernels
                            * (kh*kw each); the result equals G separate reference calls.  Used for per-image
                            * PSFs / noise levels (G = B) and for the parallel branches of net_build.jl:113-128
                            * (SHARED_INPUT | CHANNEL_CONCAT).  admmtv_backward then returns G kernels / scalars
                            * and, with SHARED_INPUT, ybar summed over the groups. */
} admmtv_desc;

int admmtv_version(void);
const char* admmtv_strerror(int code);

/* Validates the descriptor without touching the GPU. */
int admmtv_check(const admmtv_desc* desc);

/* Bytes the caller must provide: `fwd_bytes` workspace for admmtv_forward, `ckpt_bytes` for the
 * per-iteration checkpoint a later admmtv_backward needs (pass ckpt=NULL to forward for
 * inference), `bwd_bytes` workspace for admmtv_backward.  Any out pointer may be NULL. */
int admmtv_workspace_bytes(const admmtv_desc* desc, size_t* fwd_bytes, size_t* ckpt_bytes, size_t* bwd_bytes);

/* Forward.  lambda, rho (1 float each) and h (kh*kw floats) are IN/OUT: unless
 * ADMMTV_FLAG_NO_CLAMP they are clamped in place (λ,ρ to [creg,∞), h to [0,1]) exactly as the
 * reference persists the clamp into the layer struct (deconv_admm.jl:216-219).
 * bias: 1 float or NULL.  x_out: (M,N,P,B), fully overwritten.  ckpt: NULL or ckpt_bytes. */
int admmtv_forward(const admmtv_desc* desc, const float* y, float* h, float* lambda, float* rho,
                   const float* bias, float* x_out, void* workspace, void* ckpt, void* stream);

/* Backward (pullback).  xbar is the cotangent of the layer output; x_out the forward result
 * (needed for σ'); ckpt the buffer the forward call filled; y, h, lambda, rho as left by that
 * forward call (clamped).  Outputs are fully overwritten: ybar (M,N,P,B), hbar (kh*kw, may be NULL
 * when kh=0), lambdabar, rhobar (1 float each), biasbar (1 float or NULL). */
int admmtv_backward(const admmtv_desc* desc, const float* xbar, const float* x_out, const float* y,
                    const float* h, const float* lambda, const float* rho, const void* ckpt,
                    float* ybar, float* hbar, float* lambdabar, float* rhobar, float* biasbar,
                    void* workspace, void* stream);

/* Backward with the mean-squared-error loss fused in (train.jl:51-53 with loss = mean((m(x) - target)^2)): the cotangent
 * xbar = 2 (x_out - target) / numel is formed inside the first kernel and loss_sum[0] = sum((x_out - target)^2) (fp64, device;
 * the caller divides by numel) -- no xbar array and no separate loss kernel.  Everything else as admmtv_backward. */
int admmtv_backward_mse(const admmtv_desc* desc, const float* target, const float* x_out, const float* y,
                        const float* h, const float* lambda, const float* rho, const void* ckpt,
                        float* ybar, float* hbar, float* lambdabar, float* rhobar, float* biasbar,
                        double* loss_sum, void* workspace, void* stream);

/* ---- cross-rank isotropic coupling (EXTENSION) --------------------------------------
 * The reference's isotropic norm spans every image of the call (ops.jl:6,10).  When the batch is sharded over
 * several GPUs, admmtv_forward / admmtv_backward use each shard's own norm.  The _ex entry points restore the
 * single-device semantics: after the kernel that accumulates the per-pixel sum of squares (forward) or the
 * per-pixel inner products (backward) they call `allreduce_sum` on that (groups x N x M) float buffer, once per
 * iteration; the caller implements it with its communicator (ncclAllReduce on `stream`, or torch.distributed).
 * It must be stream-ordered on `stream` and return 0 on success (any other value aborts the call and is returned).
 * `tau_owner`: the per-pixel part of the lambda / rho cotangents is identical on every rank; exactly one rank
 * (tau_owner = 1) adds it, so that the usual sum all-reduce of the parameter gradients stays correct.
 * hooks = NULL behaves like admmtv_forward / admmtv_backward. */
typedef int (*admmtv_allreduce_fn)(float* buf, size_t count, void* stream, void* user);
typedef struct admmtv_hooks {
  admmtv_allreduce_fn allreduce_sum;
  void* user;
  int32_t tau_owner;
} admmtv_hooks;

int admmtv_forward_ex(const admmtv_desc* desc, const float* y, float* h, float* lambda, float* rho,
                      const float* bias, float* x_out, void* workspace, void* ckpt, void* stream,
                      const admmtv_hooks* hooks);
int admmtv_backward_ex(const admmtv_desc* desc, const float* xbar, const float* x_out, const float* y,
                       const float* h, const float* lambda, const float* rho, const void* ckpt,
                       float* ybar, float* hbar, float* lambdabar, float* rhobar, float* biasbar,
                       void* workspace, void* stream, const admmtv_hooks* hooks);

/* Host-buffer convenience (the reference's tvd_fft on a CPU Array): allocates device memory,
 * copies y/h/λ/ρ/bias in, runs admmtv_forward, copies x_out (and the clamped h/λ/ρ) back,
 * frees, synchronises.  All pointers are HOST pointers. */
int admmtv_forward_host(const admmtv_desc* desc, const float* y, float* h, float* lambda, float* rho,
                        const float* bias, float* x_out);

/* Profiling twin of admmtv_forward: same work, but brackets every launch with CUDA events on
 * `stream`, SYNCHRONISES, and returns milliseconds in ms_out[4] = {total, sum of the iterations'
 * dim-2 kernels, sum of the iterations' dim-1 kernels, everything else}.  bench.py uses it for
 * the per-kernel roofline; it is not part of the drop-in surface. */
int admmtv_profile_forward(const admmtv_desc* desc, const float* y, float* h, float* lambda, float* rho,
                           const float* bias, float* x_out, void* workspace, void* ckpt, void* stream,
                           float* ms_out);

/* Profiling twin of admmtv_backward; ms_out[4] = {total, the iterations' dim-2 kernels, the
 * iterations' dim-1 kernels, everything else}.  SYNCHRONISES. */
int admmtv_profile_backward(const admmtv_desc* desc, const float* xbar, const float* x_out, const float* y,
                            const float* h, const float* lambda, const float* rho, const void* ckpt,
                            float* ybar, float* hbar, float* lambdabar, float* rhobar, float* biasbar,
                            void* workspace, void* stream, float* ms_out);

/* Byte offsets inside the checkpoint buffer (test / inspection use): out[0] = clamp masks
 * (2 + kh*kw floats), out[1] = v_k slots, k = 1..iters-1, each float2 [Q][2][N][M] (Q = ceil(P*B/2)
 * plane pairs: .x = plane 2q, .y = plane 2q+1; [2] = dim-2 / dim-1 difference channel),
 * out[2] = F r_k slots, k = 1..iters, each float2 [Q][N][M], out[3] = isotropic per-pixel norms. */
int admmtv_ckpt_layout(const admmtv_desc* desc, size_t out[4]);

/* Number of kernel launches one admmtv_forward / admmtv_backward call enqueues (bench bookkeeping). */
int admmtv_forward_launches(const admmtv_desc* desc, int with_ckpt);
int admmtv_backward_launches(const admmtv_desc* desc);

#ifdef __cplusplus
}
#endif
#endif /* ADMMTV_H */
