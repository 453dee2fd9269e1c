/*
 * admmtv_loss.h -- C ABI of the two losses next to the ADMM-TV path.
 *
 * They produce the cotangent `xbar` that admmtv_backward consumes, so a training step stays on the
 * hand-written sm_100a kernels from the layer input to the parameter gradients:
 *
 *   reference interface replaced                                   entry point here
 *   -------------------------------------------------------------  -------------------------
 *   gmsd(x, y, t=0.0026f0, α=0f0, reduction=mean) / gmsd_loss        admmtv_gmsd_forward
 *     src/metrics/gmsd.jl:13-30, imgrads / gradientsmag
 *     src/metrics/iqa_utils.jl:24-55   (the loss of train.jl:191)
 *   Zygote pullback of it w.r.t. the FIRST argument (the prediction)   admmtv_gmsd_backward
 *   ssim(x, y, kernel; peakval=1, crop=true, dims=:) / ssim_loss /   admmtv_ssim_forward
 *     ssim_loss_fast   src/metrics/ssim.jl:84-164  (train_v2.jl:89)
 *   Zygote pullback of it w.r.t. the FIRST argument                    admmtv_ssim_backward
 *
 * Conventions are those of admmtv.h: Julia (M,N,C,B) column-major fp32 DEVICE arrays, `stream` is a
 * cudaStream_t, calls only enqueue, caller-owned 256-byte aligned workspace, 0 / <0 / >0 return codes
 * (admmtv_strerror).  Any image size is accepted (no FFT on this path).  The scalar result and its
 * cotangent are 1-float DEVICE buffers, so a training step needs no host synchronisation.
 * Only `crop=true`, `dims=:` (the defaults, the only forms the reference calls) are implemented, and
 * only separable windows (the 11-tap Gaussian of ssim.jl:6-17 and the box of ssim_loss_fast).
 */
#ifndef ADMMTV_LOSS_H
#define ADMMTV_LOSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMMTV_SSIM_MAX_TAPS 11

/* ---- GMSD ----------------------------------------------------------------------------------- */
int admmtv_gmsd_workspace_bytes(int M, int N, int C, int B, size_t* bytes);

/* loss_out[0] = mean_b sqrt(mean_{i,j,c} (gms - mean gms)^2)   (gmsd.jl:21-26).  Fills `workspace` with the
 * per-image statistics admmtv_gmsd_backward needs. */
int admmtv_gmsd_forward(int M, int N, int C, int B, int device, const float* x, const float* y, float t,
                        float alpha, float* loss_out, void* workspace, void* stream);

/* xbar (M,N,C,B), fully overwritten = lossbar[0] * d gmsd / d x.  `workspace` as left by the forward call
 * on the same x, y.  (An image whose score is exactly 0 gives non-finite values, as Zygote's sqrt rule does.) */
int admmtv_gmsd_backward(int M, int N, int C, int B, int device, const float* x, const float* y, float t,
                         float alpha, const float* lossbar, const void* workspace, float* xbar, void* stream);

/* ---- SSIM ----------------------------------------------------------------------------------- */
/* taps: HOST pointer to the L <= 11 taps of the separable window (kernel = taps * taps'), or NULL for the
 * 11-tap sigma=1.5 Gaussian of ssim.jl:6-17 (then L is ignored).  ssim_loss_fast(kernel_length=5) is
 * taps = {0.2,0.2,0.2,0.2,0.2}.  with_grad != 0 makes the forward store the three derivative maps the
 * backward reads (12 bytes per output pixel of workspace). */
int admmtv_ssim_workspace_bytes(int M, int N, int C, int B, int L, int with_grad, size_t* bytes);

/* out[0] = ssim (as_loss = 0, ssim.jl:84-124) or 1 - ssim (as_loss = 1, ssim.jl:148). */
int admmtv_ssim_forward(int M, int N, int C, int B, int device, const float* x, const float* y,
                        const float* taps, int L, float peakval, int as_loss, float* out, void* workspace,
                        int with_grad, void* stream);

/* xbar (M,N,C,B), fully overwritten = outbar[0] * d out / d x; workspace from a with_grad forward call. */
int admmtv_ssim_backward(int M, int N, int C, int B, int device, const float* x, const float* y,
                         const float* taps, int L, int as_loss, const float* outbar, const void* workspace,
                         float* xbar, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADMMTV_LOSS_H */
