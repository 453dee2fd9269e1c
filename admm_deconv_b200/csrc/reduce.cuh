// reduce.cuh -- warp-shuffle + shared-memory block reduction (fp64 scalar accumulators).
#pragma once
#include "compat.cuh"

namespace admmtv {

ADMMTV_DI double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the block; result valid in thread 0.  All threads must call.
ADMMTV_DI double block_sum(double v) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace admmtv
