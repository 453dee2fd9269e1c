// inst_small.cu -- instantiations and launcher of k_small (kernels_small.cuh): whole-call persistent kernel for plane
// pairs that fit one SM's shared memory.
#include "kernels_small.cuh"

namespace admmtv {

template <int LM, int LN>
static int launch_small(const Geom& g, const SmallArgs& a, cudaStream_t st) {
  using Cfg = SmallCfg<LM, LN>;
  static_assert(Cfg::OK, "k_small: unsupported plane size");
  if (Cfg::SMEM > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_small<LM, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return (int)e;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const int per_sm = (int)((200 * 1024) / Cfg::SMEM) < 1 ? 1 : (int)((200 * 1024) / Cfg::SMEM);
  const int slots = sms * (per_sm > 4 ? 4 : per_sm);
  const unsigned grid = (unsigned)(g.Q < slots ? g.Q : slots);
  ADMMTV_LAUNCH((k_small<LM, LN>), dim3(grid), dim3(Cfg::NT), Cfg::SMEM, st, a);
  ADMMTV_CHECK_LAUNCH();
  return 0;
}

bool small_supported(const Geom& g) {
  return (g.LM == 7 && g.LN == 7) || (g.LM == 6 && g.LN == 6) || (g.LM == 5 && g.LN == 5);
}

int run_small(const Geom& g, const SmallArgs& a, cudaStream_t st) {
  if (g.LM == 7 && g.LN == 7) return launch_small<7, 7>(g, a, st);
  if (g.LM == 6 && g.LN == 6) return launch_small<6, 6>(g, a, st);
  if (g.LM == 5 && g.LN == 5) return launch_small<5, 5>(g, a, st);
  return -3;
}

}  // namespace admmtv
