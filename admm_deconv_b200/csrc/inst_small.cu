// inst_small.cu -- instantiations and launcher of k_small (kernels_small.cuh): whole-call persistent kernel for plane
// pairs that fit one SM's shared memory.
#include "kernels_small.cuh"

namespace admmtv {

template <int LM, int LN, bool STR>
static int launch_small(const Geom& g, const SmallArgs& a, cudaStream_t st) {
  using Cfg = SmallCfg<LM, LN>;
  static_assert(Cfg::OK, "k_small: unsupported plane size");
  if (Cfg::SMEM + 2048 > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_small<LM, LN, STR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
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
  const unsigned grid = (unsigned)(a.Q < slots ? a.Q : slots);
  ADMMTV_LAUNCH((k_small<LM, LN, STR>), dim3(grid), dim3(Cfg::NT), Cfg::SMEM, st, a);
  ADMMTV_CHECK_LAUNCH();
  return 0;
}

bool small_supported(const Geom& g) {
  return (g.LM == 7 && g.LN == 7) || (g.LM == 6 && g.LN == 6) || (g.LM == 5 && g.LN == 5);
}

int run_small(const Geom& g, const SmallArgs& a_in, bool straddle, cudaStream_t st) {
  SmallArgs a = a_in;
  if (straddle) {
    // single-plane groups: the padded pairs' H^T y ([G] pairs with an empty second plane) -> [ceil(G/2)] straddling pairs,
    // written behind the state buffers' first half (v0 holds 2 * pk float2, the straddling state needs half of it)
    const int Q2 = (g.G + 1) / 2;
    float2* b2 = a.v1 + (size_t)Q2 * 2 * g.plane;   // v1 has room for g.Q = G pairs of state; the straddling kernel uses Q2 of them
    ADMMTV_LAUNCH(k_small_repack, dim3((unsigned)((g.plane + 255) / 256), (unsigned)Q2), dim3(256), 0, st, a.bpk, b2, g.plane, g.G);
    ADMMTV_CHECK_LAUNCH();
    a.bpk = b2; a.Q = Q2; a.G = g.G;
    if (g.LM == 7 && g.LN == 7) return launch_small<7, 7, true>(g, a, st);
    if (g.LM == 6 && g.LN == 6) return launch_small<6, 6, true>(g, a, st);
    if (g.LM == 5 && g.LN == 5) return launch_small<5, 5, true>(g, a, st);
    return -3;
  }
  if (g.LM == 7 && g.LN == 7) return launch_small<7, 7, false>(g, a, st);
  if (g.LM == 6 && g.LN == 6) return launch_small<6, 6, false>(g, a, st);
  if (g.LM == 5 && g.LN == 5) return launch_small<5, 5, false>(g, a, st);
  return -3;
}

}  // namespace admmtv
