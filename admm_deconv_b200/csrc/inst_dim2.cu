// inst_dim2.cu -- instantiates every dim-2 kernel variant for ONE FFT length N = 2^ADMMTV_INST.
#ifdef ADMMTV_STUB
#include "args.cuh"
namespace admmtv {
template <> int Dim2Launch<ADMMTV_INST>::run(const Geom&, int, const Dim2Args&, cudaStream_t) { return -3; }
template <> int Dim2Launch<ADMMTV_INST>::row_tile() { return 0; }
}
#else
#include "kernels.cuh"
#include "kernels_tma.cuh"
#include "kernels_cluster.cuh"

#ifndef ADMMTV_INST
#error "compile with -DADMMTV_INST=<log2 N>"
#endif
// lengths whose dim-2 pass runs the TMA-pipelined kernel k_dim2t (kernels_tma.cuh): bit (LN - 5) of this mask
#ifndef ADMMTV_D2_TMA_MASK
#define ADMMTV_D2_TMA_MASK (1 << (9 - 5))
#endif
// ... and which variants take it there: 0 = only the checkpoint-saving ones (training forward), 1 = all without accumulation
#ifndef ADMMTV_D2_TMA_ALL
#define ADMMTV_D2_TMA_ALL 0
#endif

// lengths whose inference dim-2 pass (no checkpoint, no gradient accumulation) runs the 2-CTA-cluster kernel k_dim2c
// (kernels_cluster.cuh): bit (LN - 5) of this mask
#ifndef ADMMTV_D2_CLUSTER_MASK
#define ADMMTV_D2_CLUSTER_MASK (1 << (12 - 5))   // measured (profiles/r2c_experiments.md): N = 4096 293 -> 287 us; N = 2048 216 -> 221 us (off)
#endif

namespace admmtv {

template <class K, class Args>
static int launch_k(K kern, dim3 grid, int nt, size_t smem, cudaStream_t st, const Args& a) {
  if (smem + 2048 > 48 * 1024) {   // + the kernel's static shared memory (padded to the 1 KB tile alignment)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  ADMMTV_LAUNCH(kern, grid, dim3(nt), smem, st, a);
  ADMMTV_CHECK_LAUNCH();
  return 0;
}

constexpr int LN = ADMMTV_INST;
#ifndef ADMMTV_D2_PERSIST
#define ADMMTV_D2_PERSIST 0
#endif
#ifndef ADMMTV_D2_ACC_QPB
#define ADMMTV_D2_ACC_QPB 3   // pairs per block of the G-accumulating variant: 178 -> 169 us on 96 pairs of 512^2 (2: 178, 4: 180, 6: 176)
#endif
#ifndef ADMMTV_D2_BLOCKS_PER_SM
#define ADMMTV_D2_BLOCKS_PER_SM 2
#endif

template <>
int Dim2Launch<LN>::row_tile() { return Dim2Cfg<LN>::TR; }

template <>
int Dim2Launch<LN>::run(const Geom& g, int variant, const Dim2Args& a_in, cudaStream_t st) {
  using Cfg = Dim2Cfg<LN>;
  // One wave of persistent blocks: each block owns a row tile and loops over plane pairs, so the
  // gradient accumulator stays in registers and the next tile can be prefetched into L2.
  Dim2Args a = a_in;
  a.Q = g.Q;
  const int row_tiles = g.M / Cfg::TR;
  int gy = g.Q;
#if ADMMTV_D2_PERSIST
  if (ADMMTV_D2_PERSIST == 1 || variant == D2_C_ACCG) {  // 2: persistent only where it saves atomics
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int slots = sms * ADMMTV_D2_BLOCKS_PER_SM;
    gy = slots / row_tiles;
    if (gy < 1) gy = 1;
    if (gy > g.Q) gy = g.Q;
  }
#endif
  // the G-accumulating variant gives each block a few pairs so the shared-memory partial sums are flushed
  // with one global atomic per element per ADMMTV_D2_ACC_QPB pairs
  if (variant == D2_C_ACCG && gy == g.Q) gy = (g.Q + ADMMTV_D2_ACC_QPB - 1) / ADMMTV_D2_ACC_QPB;
  if (gy > 65535) gy = 65535;  // blocks loop over pairs with stride gridDim.y
#ifndef ADMMTV_EMU
  if constexpr (LN >= 10 && LN <= 12 && ((ADMMTV_D2_CLUSTER_MASK >> (LN - 5)) & 1)) {
    if constexpr (Dim2cCfg<LN>::OK) {
      using CC = Dim2cCfg<LN>;
      const dim3 cgrid((unsigned)(CC::CS * (g.M / CC::TR)), (unsigned)gy);   // x = CS * row tile + rank in the cluster
      // once per process: can a cluster of CS such blocks be co-scheduled at all on this device / partition?
      static const bool cluster_ok = [] {
        int n = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CC::CS, 1, 1); cfg.blockDim = dim3(CC::NT, 1, 1); cfg.dynamicSmemBytes = CC::SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CC::CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaFuncSetAttribute(k_dim2c<LN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CC::SMEM) != cudaSuccess ||
            cudaOccupancyMaxActiveClusters(&n, k_dim2c<LN, 0>, &cfg) != cudaSuccess) {
          cudaGetLastError();
          return false;
        }
        return n > 0;
      }();
      if (cluster_ok && g.M % CC::TR == 0) switch (variant) {
        case D2_C: return launch_k(k_dim2c<LN, 0>, cgrid, CC::NT, CC::SMEM, st, a);
        case D2_KCONJ: return launch_k(k_dim2c<LN, 1>, cgrid, CC::NT, CC::SMEM, st, a);
        default: break;   // D2_K (backward without a PSF gradient) and every saving / accumulating variant: k_dim2
      }
    }
  }
  if constexpr (LN >= 5 && LN <= 12 && Dim2tCfg<LN>::OK && ((ADMMTV_D2_TMA_MASK >> (LN - 5)) & 1)) {
    // TMA-pipelined persistent kernel for the variants without gradient accumulation
    if (variant != D2_C_ACCG && variant != D2_K_ACCP && (ADMMTV_D2_TMA_ALL || variant == D2_C_SAVE || variant == D2_KCONJ_SAVE)) {
      using T = Dim2tCfg<LN>;
      CUtensorMap map;
      if (tma_make_map(&map, a.in, g.M, g.N, g.Q, T::TR, T::BOXC) == 0) {
        static int sms = 0;
        if (sms == 0) {
          int dev = 0;
          cudaGetDevice(&dev);
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const int total = row_tiles * g.Q;
        const dim3 tgrid((unsigned)(total < sms ? total : sms));
        constexpr int TNT = T::NT;
#define ADMMTV_LAUNCH_T(KERN)                                                                                          \
  do {                                                                                                                 \
    cudaError_t e = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);             \
    if (e != cudaSuccess) return (int)e;                                                                               \
    KERN<<<tgrid, dim3(TNT), T::SMEM, st>>>(a, map);                                                                   \
    ADMMTV_CHECK_LAUNCH();                                                                                             \
    return 0;                                                                                                          \
  } while (0)
        switch (variant) {
          case D2_C: ADMMTV_LAUNCH_T((k_dim2t<LN, 0, false, false>));
          case D2_C_SAVE: ADMMTV_LAUNCH_T((k_dim2t<LN, 0, true, false>));
          case D2_KCONJ: ADMMTV_LAUNCH_T((k_dim2t<LN, 1, false, false>));
          case D2_KCONJ_SAVE: ADMMTV_LAUNCH_T((k_dim2t<LN, 1, true, false>));
          case D2_K: ADMMTV_LAUNCH_T((k_dim2t<LN, 2, false, false>));
          case D2_FWDONLY: ADMMTV_LAUNCH_T((k_dim2t<LN, 0, false, true>));
          default: break;
        }
#undef ADMMTV_LAUNCH_T
      }
    }
  }
#endif
  const dim3 grid((unsigned)row_tiles, (unsigned)gy);
  switch (variant) {
    case D2_C: return launch_k(k_dim2<LN, 0, false, 0, false>, grid, Cfg::NT, Cfg::SMEM, st, a);
    case D2_C_SAVE: return launch_k(k_dim2<LN, 0, true, 0, false>, grid, Cfg::NT, Cfg::SMEM, st, a);
    case D2_KCONJ: return launch_k(k_dim2<LN, 1, false, 0, false>, grid, Cfg::NT, Cfg::SMEM, st, a);
    case D2_C_ACCG: return launch_k(k_dim2<LN, 0, false, 1, false>, grid, Cfg::NT, Cfg::SMEM + Cfg::SMEM / 2, st, a);
    case D2_FWDONLY: return launch_k(k_dim2<LN, 0, false, 0, true>, grid, Cfg::NT, Cfg::SMEM, st, a);
    case D2_K_ACCP: return launch_k(k_dim2<LN, 2, false, 2, false>, grid, Cfg::NT, Cfg::SMEM, st, a);
    case D2_K: return launch_k(k_dim2<LN, 2, false, 0, false>, grid, Cfg::NT, Cfg::SMEM, st, a);
    case D2_KCONJ_SAVE: return launch_k(k_dim2<LN, 1, true, 0, false>, grid, Cfg::NT, Cfg::SMEM, st, a);
  }
  return -5;
}

}  // namespace admmtv
#endif  // ADMMTV_STUB
