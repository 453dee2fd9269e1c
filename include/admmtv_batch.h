/*
 * admmtv_batch.h -- C ABI of the batch-assembly step in front of the ADMM-TV path.
 *
 *   reference interface replaced                                        entry point here
 *   ------------------------------------------------------------------  -------------------------
 *   img2tensor (N0f8 channel-interleaved image -> Float32 (H,W,C))        admmtv_batch_from_n0f8
 *     src/utilities/base_funcs.jl:29-35, followed by the `cat(...; dims=4)`
 *     of the crops into the (H,W,C,B) batch, src/processing/datafeeder.jl:54-68,
 *     and the host->device copy of `|> gpu` (train.jl:50, ToGPU() train_v2.jl:60)
 *
 * The reference converts to Float32 and concatenates on the host, then uploads 4 bytes per sample.  Here the
 * host only gathers the raw 8-bit crops into one pinned buffer (1 byte per sample over PCIe); the conversion,
 * the channel de-interleave and the placement into the (M,N,C,B) batch happen in one kernel on the device.
 * Conventions as in admmtv.h (device pointers, stream-ordered, integer return codes).
 */
#ifndef ADMMTV_BATCH_H
#define ADMMTV_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dst (M,N,C,B) column-major fp32, fully overwritten:
 *   dst[i + M*(j + N*(c + C*b))] = src[b*stride_b + c*stride_c + i*stride_i + j*stride_j] / 255
 * (the N0f8 -> Float32 conversion of img2tensor).  Strides are in bytes (= elements).  A Julia
 * Matrix{RGB{N0f8}} crop gathered per image has (stride_c, stride_i, stride_j) = (1, C, C*M); a row-major
 * (H,W,C) uint8 crop (numpy / PIL) has (1, C*N, C). */
int admmtv_batch_from_n0f8(int M, int N, int C, int B, int device, const uint8_t* src, int64_t stride_c,
                           int64_t stride_i, int64_t stride_j, int64_t stride_b, float* dst, void* stream);

/* Device-resident dataset variant: the decoded 8-bit images stay in HBM (180 GB per B200 holds ~290k RGB 512x512
 * frames) and a batch is gathered from them on the device -- per step the host only sends B offsets.
 *   dst[i + M*(j + N*(c + C*b))] = base[offsets[b] + c*stride_c + i*stride_i + j*stride_j] / 255
 * `offsets` is a DEVICE array of B int64 element offsets (image start + crop origin: h0*stride_i + w0*stride_j for the
 * random aligned crops of datafeeder.jl:43-45); strides are those of the FULL images, e.g. row-major (H,W,C):
 * (1, C*W, C).  All images of one call share the strides (same width); x and y batches are two calls. */
int admmtv_batch_gather_n0f8(int M, int N, int C, int B, int device, const uint8_t* base, const int64_t* offsets,
                             int64_t stride_c, int64_t stride_i, int64_t stride_j, float* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADMMTV_BATCH_H */
