// inst_generic.cu -- launchers of the generic-size kernels (generic_kernels.cuh) as size id 0 of the
// Dim1Launch / Dim2Launch contracts (args.cuh): any M, N in 1..4096 without a register-FFT plan.
// Each contract of the tuned path (one fused kernel) becomes a short sequence here:
//   [inverse dim-1 DFT, in place] -> per-pixel sweep -> [forward dim-1 DFT, in place].
// In-place is safe: every spectrum buffer handed to a dim-1 contract is consumed by that launch only
// (admmtv_api.cu / bwd_api.inc rewrite it with the next k_dim2 before it is read again).
#include "generic_kernels.cuh"

namespace admmtv {

static int gk_set_smem(const void* kern, size_t smem) {
  if (smem + 2048 > 48 * 1024) {   // + the kernel's static shared memory (padded to the 1 KB tile alignment)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

static int gk_run_dft1(const Geom& g, const float2* in, float2* out, const float2* tw, bool inv, cudaStream_t st) {
  GDft1Args a{};
  a.in = in; a.out = out; a.tw = tw; a.M = g.M; a.P = dft_plan(g.M);
  a.LB = gk_lines_dim1(g.M); a.inv = inv ? 1 : 0; a.nlines = (long long)g.N * g.Q;
  const size_t smem = gk_dft_smem_elems(g.M, a.P, a.LB) * sizeof(float2);
  int rc = gk_set_smem((const void*)gk_dft1, smem);
  if (rc) return rc;
  const long long nblk = (a.nlines + a.LB - 1) / a.LB;
  ADMMTV_LAUNCH(gk_dft1, dim3((unsigned)nblk), dim3(GK_NT), smem, st, a);
  ADMMTV_CHECK_LAUNCH();
  return 0;
}
static unsigned gk_pixel_blocks(const Geom& g) { return (unsigned)((g.pk + GK_NT - 1) / GK_NT); }

template <>
int Dim1Launch<0>::col_tile() { return 0; }

template <>
int Dim1Launch<0>::pack_fft1(const Geom& g, int mode, const PackArgs& a, cudaStream_t st) {
  ADMMTV_LAUNCH(gk_pack, dim3(gk_pixel_blocks(g)), dim3(GK_NT), 0, st, a, g.M, g.Q, mode);
  ADMMTV_CHECK_LAUNCH();
  return gk_run_dft1(g, a.spec, a.spec, a.twM, false, st);
}
template <>
int Dim1Launch<0>::out(const Geom& g, int mode, const OutArgs& a, cudaStream_t st) {
  int rc = gk_run_dft1(g, a.spec, const_cast<float2*>(a.spec), a.twM, true, st);
  if (rc) return rc;
  ADMMTV_LAUNCH(gk_out, dim3(gk_pixel_blocks(g)), dim3(GK_NT), 0, st, a, g.M, g.Q, mode);
  ADMMTV_CHECK_LAUNCH();
  return 0;
}
static int gk_fwd(const Geom& g, const Dim1FwdArgs& a, int mode, bool has_vprev, cudaStream_t st) {
  int rc;
  if (mode != 1 && (rc = gk_run_dft1(g, a.spec_in, const_cast<float2*>(a.spec_in), a.twM, true, st))) return rc;
  ADMMTV_LAUNCH(gk_sweep_fwd, dim3(gk_pixel_blocks(g)), dim3(GK_NT), 0, st, a, a.spec_in, a.spec_out, g.M, g.Q, mode,
                has_vprev ? 1 : 0);
  ADMMTV_CHECK_LAUNCH();
  if (mode != 2 && (rc = gk_run_dft1(g, a.spec_out, a.spec_out, a.twM, false, st))) return rc;
  return 0;
}
template <>
int Dim1Launch<0>::fwd(const Geom& g, bool has_vprev, const Dim1FwdArgs& a, cudaStream_t st) { return gk_fwd(g, a, 0, has_vprev, st); }
template <>
int Dim1Launch<0>::fwd_iso_a(const Geom& g, bool has_vprev, const Dim1FwdArgs& a, cudaStream_t st) { return gk_fwd(g, a, 2, has_vprev, st); }
template <>
int Dim1Launch<0>::fwd_iso_b(const Geom& g, const Dim1FwdArgs& a, cudaStream_t st) { return gk_fwd(g, a, 1, true, st); }

static int gk_bwd(const Geom& g, const Dim1BwdArgs& a, int mode, bool has_vbar, cudaStream_t st) {
  int rc;
  // pass B of the isotropic backward reads the rbar_k that pass A already brought back to the spatial domain
  if (mode != 1 && (rc = gk_run_dft1(g, a.spec_in, const_cast<float2*>(a.spec_in), a.twM, true, st))) return rc;
  ADMMTV_LAUNCH(gk_sweep_bwd, dim3(gk_pixel_blocks(g)), dim3(GK_NT), 0, st, a, a.spec_in, a.spec_out, g.M, g.Q, mode,
                has_vbar ? 1 : 0);
  ADMMTV_CHECK_LAUNCH();
  if (mode != 2 && (rc = gk_run_dft1(g, a.spec_out, a.spec_out, a.twM, false, st))) return rc;
  return 0;
}
template <>
int Dim1Launch<0>::bwd(const Geom& g, bool has_vbar, const Dim1BwdArgs& a, cudaStream_t st) { return gk_bwd(g, a, 0, has_vbar, st); }
template <>
int Dim1Launch<0>::bwd_iso_a(const Geom& g, bool has_vbar, const Dim1BwdArgs& a, cudaStream_t st) { return gk_bwd(g, a, 2, has_vbar, st); }
template <>
int Dim1Launch<0>::bwd_iso_b(const Geom& g, bool has_vbar, const Dim1BwdArgs& a, cudaStream_t st) { return gk_bwd(g, a, 1, has_vbar, st); }
template <>
int Dim1Launch<0>::bwd_last(const Geom& g, int mode, const Dim1BwdArgs& a, cudaStream_t st) {
  int rc = gk_run_dft1(g, a.spec_in, const_cast<float2*>(a.spec_in), a.twM, true, st);
  if (rc) return rc;
  ADMMTV_LAUNCH(gk_bwd_last, dim3(gk_pixel_blocks(g)), dim3(GK_NT), 0, st, a, a.spec_in, a.spec_out, g.M, g.Q, mode);
  ADMMTV_CHECK_LAUNCH();
  if (mode == 0) return gk_run_dft1(g, a.spec_out, a.spec_out, a.twM, false, st);
  return 0;
}

template <>
int Dim2Launch<0>::row_tile() { return 0; }

template <>
int Dim2Launch<0>::run(const Geom& g, int variant, const Dim2Args& a_in, cudaStream_t st) {
  Dim2Args a = a_in;
  a.Q = g.Q;
  GDim2Cfg c{};
  c.N = g.N; c.P = dft_plan(g.N); c.TR = gk_rows_dim2(g.N);
  switch (variant) {   // (MUL, SAVE_Z, ACC, FWD_ONLY) as in inst_dim2.cu
    case D2_C: break;
    case D2_C_SAVE: c.save_z = 1; break;
    case D2_KCONJ: c.mul = 1; break;
    case D2_C_ACCG: c.acc = 1; break;
    case D2_FWDONLY: c.fwd_only = 1; break;
    case D2_K_ACCP: c.mul = 2; c.acc = 2; break;
    case D2_K: c.mul = 2; break;
    case D2_KCONJ_SAVE: c.mul = 1; c.save_z = 1; break;
    default: return -5;
  }
  const size_t smem = gk_dft_smem_elems(g.N, c.P, c.TR) * sizeof(float2);
  int rc = gk_set_smem((const void*)gk_dim2, smem);
  if (rc) return rc;
  const size_t nblk = (size_t)((g.M + c.TR - 1) / c.TR) * g.Q;
  ADMMTV_LAUNCH(gk_dim2, dim3((unsigned)nblk), dim3(GK_NT), smem, st, a, c);
  ADMMTV_CHECK_LAUNCH();
  return 0;
}

}  // namespace admmtv
