// inst_dim1.cu -- instantiates every dim-1 kernel for ONE FFT length M = 2^ADMMTV_INST.
// Compiled once per supported length (see admm_deconv_b200/build.py).
#ifdef ADMMTV_STUB
#include "args.cuh"
namespace admmtv {
constexpr int LM_ = ADMMTV_INST;
template <> int Dim1Launch<LM_>::pack_fft1(const Geom&, int, const PackArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::out(const Geom&, int, const OutArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::fwd(const Geom&, bool, const Dim1FwdArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::bwd(const Geom&, bool, const Dim1BwdArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::bwd_last(const Geom&, int, const Dim1BwdArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::fwd_iso_b(const Geom&, const Dim1FwdArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::fwd_iso_a(const Geom&, bool, const Dim1FwdArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::bwd_iso_b(const Geom&, bool, const Dim1BwdArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::bwd_iso_a(const Geom&, bool, const Dim1BwdArgs&, cudaStream_t) { return -3; }
template <> int Dim1Launch<LM_>::col_tile() { return 0; }
}
#else
#include "kernels.cuh"
#include "kernels_bwd.cuh"

#ifndef ADMMTV_INST
#error "compile with -DADMMTV_INST=<log2 M>"
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

constexpr int LM = ADMMTV_INST;
using Cfg = Dim1Cfg<LM>;
static dim3 dim1_grid(const Geom& g, int co = Cfg::CO) { return dim3((unsigned)((g.N + co - 1) / co) * (unsigned)g.Q); }

// TMA variants of the iteration kernels: launched when the column maps of the two spectra can be encoded (a null pointer =
// the kernel does not touch that side), else the caller falls through to the LDG / STG kernels.
#ifndef ADMMTV_EMU
template <class K, class Args>
static int launch_tma(K kern, const Geom& g, int nt, const float2* in, const float2* out, cudaStream_t st, const Args& a, int co = Cfg::CO,
                      size_t smem = Cfg::SMEM) {
  Dim1Tma tm;
  const size_t ncols = (size_t)g.Q * g.N;
  if (g.N % co != 0 || ncols >= 0x7fffffffull) return -100;
  if (tma_make_colmap(&tm.in, in ? in : out, g.M, ncols) != 0 || tma_make_colmap(&tm.out, out ? out : in, g.M, ncols) != 0) return -100;
  if (smem + 2048 > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<dim1_grid(g, co), dim3(nt), smem, st>>>(a, tm);
  ADMMTV_CHECK_LAUNCH();
  return 0;
}
#define ADMMTV_TRY_TMA(KERN, NT_, IN, OUT, ...)                                 \
  if constexpr (kDim1TmaOk<LM>) {                                               \
    const int rc__ = launch_tma(KERN, g, NT_, IN, OUT, st, a, ##__VA_ARGS__);   \
    if (rc__ != -100) return rc__;                                              \
  }
#else
#define ADMMTV_TRY_TMA(KERN, NT_, IN, OUT, ...)
#endif

template <>
int Dim1Launch<LM>::col_tile() { return Cfg::CO; }

template <>
int Dim1Launch<LM>::pack_fft1(const Geom& g, int mode, const PackArgs& a, cudaStream_t st) {
  switch (mode) {
    case 0: { ADMMTV_TRY_TMA((k_pack_fft1_tma<LM, 0>), Cfg::NT, (const float2*)nullptr, a.spec); break; }
    case 1: { ADMMTV_TRY_TMA((k_pack_fft1_tma<LM, 1>), Cfg::NT, (const float2*)nullptr, a.spec); break; }
    case 3: { ADMMTV_TRY_TMA((k_pack_fft1_tma<LM, 3>), Cfg::NT, (const float2*)nullptr, a.spec); break; }
    default: { ADMMTV_TRY_TMA((k_pack_fft1_tma<LM, 2>), Cfg::NT, (const float2*)nullptr, a.spec); break; }
  }
  if (mode == 0) return launch_k(k_pack_fft1<LM, 0>, dim1_grid(g), Cfg::NT, Cfg::SMEM, st, a);
  if (mode == 1) return launch_k(k_pack_fft1<LM, 1>, dim1_grid(g), Cfg::NT, Cfg::SMEM, st, a);
  if (mode == 3) return launch_k(k_pack_fft1<LM, 3>, dim1_grid(g), Cfg::NT, Cfg::SMEM, st, a);
  return launch_k(k_pack_fft1<LM, 2>, dim1_grid(g), Cfg::NT, Cfg::SMEM, st, a);
}
template <>
int Dim1Launch<LM>::out(const Geom& g, int mode, const OutArgs& a, cudaStream_t st) {
  switch (mode) {
    case 0: { ADMMTV_TRY_TMA((k_dim1_out_tma<LM, 0>), Cfg::NT, a.spec, (const float2*)nullptr); break; }
    case 2: { ADMMTV_TRY_TMA((k_dim1_out_tma<LM, 2>), Cfg::NT, a.spec, (const float2*)nullptr); break; }
    default: { ADMMTV_TRY_TMA((k_dim1_out_tma<LM, 1>), Cfg::NT, a.spec, (const float2*)nullptr); break; }
  }
  if (mode == 0) return launch_k(k_dim1_out<LM, 0>, dim1_grid(g), Cfg::NT, Cfg::SMEM, st, a);
  if (mode == 2) return launch_k(k_dim1_out<LM, 2>, dim1_grid(g), Cfg::NT, Cfg::SMEM, st, a);
  return launch_k(k_dim1_out<LM, 1>, dim1_grid(g), Cfg::NT, Cfg::SMEM, st, a);
}
template <>
int Dim1Launch<LM>::fwd(const Geom& g, bool has_vprev, const Dim1FwdArgs& a, cudaStream_t st) {
  ADMMTV_TRY_TMA((has_vprev ? k_dim1_fwd_tma<LM, true> : k_dim1_fwd_tma<LM, false>), (Dim1FwdCfg<LM, 0>::NT), a.spec_in, a.spec_out);
  if (has_vprev) return launch_k(k_dim1_fwd<LM, true>, dim1_grid(g), Dim1FwdCfg<LM, 0>::NT, Cfg::SMEM, st, a);
  return launch_k(k_dim1_fwd<LM, false>, dim1_grid(g), Dim1FwdCfg<LM, 0>::NT, Cfg::SMEM, st, a);
}
#include "inst_dim1_bwd.inc"

}  // namespace admmtv
#endif  // ADMMTV_STUB
