/* admmtv_host.h -- HOST-BUFFER entry points of libadmmtv.so: the calls a host-side caller makes when its arrays live
 * in CPU memory.
 *
 * Reference interfaces these replace:
 *   - tvd_fft on a CPU Array                        /root/reference/src/ops/ops.jl:183-187  (-> tvd_fft_cpu :17-96)
 *   - one training step of train.jl                 /root/reference/src/train.jl:49-54: batch from the DataLoader (host
 *     Float32 (M,N,P,B) arrays, processing/datafeeder.jl:54-68) |> gpu, Flux.withgradient(model) do loss(m(x), y) end
 *
 * They still run on the GPU (there is no CPU fallback): the session keeps the device buffers, two input / output slots
 * and three streams, so that step i+1's host->device copy and step i-1's device->host copy overlap step i's kernels.
 * All `const float*` / `float*` arguments below are HOST pointers in the Julia (M,N,P,B) column-major layout of
 * admmtv.h.  Pin them (admmtv_host_pin, or your runtime's pinned allocator) to get asynchronous full-speed PCIe
 * transfers; pageable memory works but every copy then blocks the calling thread.
 */
#ifndef ADMMTV_HOST_H
#define ADMMTV_HOST_H

#include "admmtv.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct admmtv_host_session admmtv_host_session;

/* Device bytes a session needs (inputs/outputs of both slots, workspaces, and for training = 1 the per-iteration
 * checkpoint and the backward workspace). */
int admmtv_host_session_bytes(const admmtv_desc* desc, int training, size_t* device_bytes);

/* device_arena: NULL (the library cudaMallocs / frees it) or a caller-owned 256-byte aligned device buffer of
 * admmtv_host_session_bytes bytes (e.g. a CuArray{UInt8}, so the caller's pool accounts for it).
 * compute_stream: NULL (the session creates one) or the caller's stream -- the kernels and the optional gradient
 * all-reduce hook are enqueued on it.  The copy streams are always the session's own. */
int admmtv_host_session_create(const admmtv_desc* desc, int training, void* device_arena, void* compute_stream,
                               admmtv_host_session** out);
int admmtv_host_session_destroy(admmtv_host_session* s);

/* cudaHostRegister / cudaHostUnregister of a caller-owned host range. */
int admmtv_host_pin(void* host_ptr, size_t bytes);
int admmtv_host_unpin(void* host_ptr);

/* Inference through slot 0 or 1: y -> device, admmtv_forward, x -> x_out.  h / lambda / rho (host, IN/OUT: the
 * persisted clamp of deconv_admm.jl:216-219 is written back), bias as in admmtv_forward.  Returns after ENQUEUEING;
 * admmtv_host_wait(slot) blocks until x_out (and the clamped parameters) are in host memory.  A slot must be waited
 * for before it is enqueued again. */
int admmtv_host_forward_enqueue(admmtv_host_session* s, int slot, const float* y, float* h, float* lambda, float* rho,
                                const float* bias, float* x_out);

/* The same forward call fed with 8-bit samples (N0f8, value / 255: a decoded image before `img2tensor`, base_funcs.jl:29-35;
 * admm_deconv_test.jl:60-76 converts on the CPU and uploads fp32): `y` is a HOST byte array with ELEMENT strides
 * stride_c / stride_i / stride_j / stride_b as in admmtv_batch_from_n0f8; one byte per sample crosses PCIe and the fp32
 * batch is built on the device.  x_out is fp32 as above.  Bit-identical to the float call on value / 255. */
int admmtv_host_forward_enqueue_n0f8(admmtv_host_session* s, int slot, const uint8_t* y, int64_t stride_c, int64_t stride_i,
                                     int64_t stride_j, int64_t stride_b, float* h, float* lambda, float* rho, const float* bias,
                                     float* x_out);

/* One training step through slot 0 or 1 (train.jl:49-54 with the mean-squared-error loss):
 *   y, target -> device ; x = layer(y) with the per-iteration checkpoint ; loss = mean((x - target)^2),
 *   xbar = 2 (x - target) / numel ; admmtv_backward ; [hooks->allreduce_sum on the packed gradient buffer, once:
 *   the data-parallel gradient all-reduce over NVLink] ; gradients -> host.
 * grads_out (host, fully overwritten) = [hbar (kh*kw*G) | lambdabar (G) | rhobar (G) | biasbar (G, only with has_bias)]
 * with G = max(groups, 1): admmtv_host_grad_floats(desc) floats.  loss_out: 1 float (this rank's shard), written by
 * admmtv_host_wait.  ybar_out: NULL, or (M,N,P,B) floats for the cotangent of the input (not needed when the layer is
 * the first of the net).  hooks may be NULL. */
int admmtv_host_train_step_enqueue(admmtv_host_session* s, int slot, const float* y, const float* target, float* h,
                                   float* lambda, float* rho, const float* bias, float* grads_out, float* loss_out,
                                   float* ybar_out, const admmtv_hooks* hooks);

/* The same step fed with the dataset's own sample format: `y` and `target` are HOST arrays of 8-bit samples (N0f8, value / 255:
 * what `load(img)` yields before `img2tensor`, base_funcs.jl:29-35) of the batch's (M,N,P,B) images with ELEMENT strides
 * stride_c / stride_i / stride_j / stride_b (channel, dim 1, dim 2, image) as in admmtv_batch_from_n0f8 (admmtv_batch.h).
 * One byte per sample crosses PCIe instead of four; conversion, de-interleave and placement run on the device (the kernel
 * of admmtv_batch_from_n0f8) in front of the forward.  Replaces datafeeder.jl:54-68 + `|> gpu` (train.jl:50) + the step.
 * Bit-identical to admmtv_host_train_step_enqueue on the float arrays value / 255.  No cotangent of the input is returned. */
int admmtv_host_train_step_enqueue_n0f8(admmtv_host_session* s, int slot, const uint8_t* y, const uint8_t* target,
                                        int64_t stride_c, int64_t stride_i, int64_t stride_j, int64_t stride_b, float* h,
                                        float* lambda, float* rho, const float* bias, float* grads_out, float* loss_out,
                                        const admmtv_hooks* hooks);
int admmtv_host_grad_floats(const admmtv_desc* desc);

/* Device-pointer twin of the training step above (inputs already resident in HBM; every pointer is a DEVICE pointer,
 * caller-owned, stream-ordered on `stream`, returns after enqueueing): admmtv_forward with checkpoint ->
 * admmtv_backward_mse (the pullback seed 2 (x - target) / numel is formed inside its first kernel; loss_sum[0] =
 * sum((x - target)^2) in fp64, the caller divides by numel) -> [hooks->allreduce_sum(grads_packed)].
 * ybar: (M,N,P,B) cotangent of y; grads_packed as grads_out above.  This is the step bench.py times with resident
 * inputs; the host session calls exactly this. */
int admmtv_mse_train_step(const admmtv_desc* desc, const float* y, const float* target, float* h, float* lambda,
                          float* rho, const float* bias, float* x_out, float* ybar, float* grads_packed,
                          double* loss_sum, void* ws_fwd, void* ckpt, void* ws_bwd, void* stream,
                          const admmtv_hooks* hooks);

/* Blocks until everything enqueued through `slot` has completed and its results are in host memory. */
int admmtv_host_wait(admmtv_host_session* s, int slot);

/* Kernel launches one train step / forward of this session enqueues (bench bookkeeping). */
int admmtv_host_launches(const admmtv_host_session* s, int training);

#ifdef __cplusplus
}
#endif
#endif /* ADMMTV_HOST_H */
