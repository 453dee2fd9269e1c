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
 * admmtv_ssim_forward / _backward take a separable window with equal taps in both dimensions (the 11-tap
 * Gaussian of ssim.jl:6-17, the box of ssim_loss_fast: the forms the reference's scripts use, fully unrolled
 * kernels); admmtv_ssim_window_* take ANY (L1, L2) window `kernel_ref` of ssim.jl:84 as a sum of R separable
 * terms (the host side factors it, e.g. by an SVD); `crop=false` (ssim.jl:104-110) is admmtv_pad_symmetric on
 * both images followed by the valid-size call, its pullback admmtv_pad_symmetric_adjoint.  `dims` is accepted
 * and ignored by the reference (ssim.jl:84-124 never reads it).  Both losses are symmetric in (x, y), so the
 * pullback w.r.t. the SECOND argument is the same entry points with the two images swapped.
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

/* ---- SSIM with an arbitrary window (ssim.jl:84 `kernel_ref`, any (L1, L2) <= 11 x 11 array) ------------- */
/* The window is given as R <= 11 separable terms, W[a,b] = sum_r u[r*L1 + a] * v[r*L2 + b] (HOST pointers; a
 * rank-1 window has R = 1, an SVD gives R = min(L1, L2) for any other).  Output map (M-L1+1, N-L2+1). */
int admmtv_ssim_window_workspace_bytes(int M, int N, int C, int B, int L1, int L2, int with_grad, size_t* bytes);

int admmtv_ssim_window_forward(int M, int N, int C, int B, int device, const float* x, const float* y,
                               const float* u, const float* v, int L1, int L2, int R, float peakval, int as_loss,
                               float* out, void* workspace, int with_grad, void* stream);

int admmtv_ssim_window_backward(int M, int N, int C, int B, int device, const float* x, const float* y,
                                const float* u, const float* v, int L1, int L2, int R, int as_loss,
                                const float* outbar, const void* workspace, float* xbar, void* stream);

/* ---- pad_symmetric (NNlib; ssim.jl:104-110, `crop = false`) --------------------------------------------- */
/* dst (M+lo1+hi1, N+lo2+hi2, planes) = src (M, N, planes) mirrored across its borders INCLUDING the edge sample;
 * ssim.jl:107 pads by (cld(L-1,2), fld(L-1,2)) per dimension.  Pads may not exceed the array size (NNlib). */
int admmtv_pad_symmetric(int M, int N, int planes, int lo1, int hi1, int lo2, int hi2, int device, const float* src,
                         float* dst, void* stream);

/* src_bar (M, N, planes), fully overwritten = the pullback of the padding applied to padded_bar. */
int admmtv_pad_symmetric_adjoint(int M, int N, int planes, int lo1, int hi1, int lo2, int hi2, int device,
                                 const float* padded_bar, float* src_bar, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADMMTV_LOSS_H */
