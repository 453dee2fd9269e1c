// fft_core.cuh -- register-resident radix-{2,4,8,16} butterflies, the per-length stage plans
// and the digit-reversed position<->frequency maps shared by every kernel of the path.
//
// Design (DESIGN.md "FFT"): a length-L complex transform is a sequence of NS in-place passes.
// The forward transform is decimation-in-frequency (natural order in, digit-reversed order
// out); the inverse is the exact transpose, decimation-in-time (digit-reversed in, natural
// out).  Because every spectral operation of the ADMM iteration is a pointwise multiply by a
// table, the tables are stored in digit-reversed order and no reordering pass ever runs.
//
// Replaces: CUFFT.rfft / CUFFT.irfft at /root/reference/src/ops/ops.jl:108,117-118,168 and
// FFTW rfft/irfft at :26,35-36,86.
#pragma once

#include "compat.cuh"

namespace admmtv {

// ------------------------------------------------------------------------------------------
// stage plans (DIF order: stage 0 works on the whole line, the last stage on R-long blocks)
// ------------------------------------------------------------------------------------------
// Plan ids: 0 = the plan of the dim-1 (contiguous) transforms, 1 = the plan of the dim-2 (strided) pass.  They differ
// where the dim-2 kernels profit from fewer, larger passes: those kernels are bound by the shared-memory / L1 data path
// (ncu: l1tex 73 % vs DRAM 45 %), every pass costs one read and one write of the tile, and their work-item mapping
// (rows x N / R items for 256 threads) fits any radix.  512 = 16 x 32 makes their forward + inverse transform 3 passes
// instead of 5.  The spectrum order along a dimension follows the plan of that dimension (pos_to_freq).
#ifndef ADMMTV_PLAN2_512
#define ADMMTV_PLAN2_512 0   // measured on B200 (64 x 512^2 x 3): 16 x 32 needs 128 registers (2 blocks/SM): k_dim2 97 -> 103 us, save variant 135 -> 120, accG 179 -> 192: net zero, off
#endif
#ifndef ADMMTV_PLAN2_2048
#define ADMMTV_PLAN2_2048 0   // dim-2 plan of 2048: 0 = 16 x 16 x 8 (as dim 1), 1 = 8 x 8 x 8 x 4, 2 = 4 x 8 x 8 x 8 (radix-8 passes: fewer registers, more threads)
#endif
#ifndef ADMMTV_PLANC_1024
#define ADMMTV_PLANC_1024 0   // plan id 2 = the half-line transforms of the 2-CTA-cluster dim-2 pass (kernels_cluster.cuh); 1024: 0 = 16 x 8 x 8, 1 = 8 x 8 x 16, 2 = 8 x 16 x 8
#endif
ADMMTV_HD constexpr int plan_radix(int L, int s, int P = 0) {
  if (P == 2 && ADMMTV_PLANC_1024 == 1 && L == 1024) return s < 2 ? 8 : (s == 2 ? 16 : 1);
  if (P == 2 && ADMMTV_PLANC_1024 == 2 && L == 1024) return s == 0 ? 8 : (s == 1 ? 16 : (s == 2 ? 8 : 1));
  if (P == 1 && ADMMTV_PLAN2_512 && L == 512) return s == 0 ? 16 : (s == 1 ? 32 : 1);
  if (P == 1 && ADMMTV_PLAN2_2048 == 1 && L == 2048) return s < 3 ? 8 : (s == 3 ? 4 : 1);
  if (P == 1 && ADMMTV_PLAN2_2048 == 2 && L == 2048) return s == 0 ? 4 : (s < 4 ? 8 : 1);
  switch (L) {
    case 32:   return s == 0 ? 8 : (s == 1 ? 4 : 1);
    case 64:   return s < 2 ? 8 : 1;
    case 128:  return s == 0 ? 16 : (s == 1 ? 8 : 1);
    case 256:  return s < 2 ? 16 : 1;
    case 512:  return s < 3 ? 8 : 1;
    case 1024: return s == 0 ? 16 : (s < 3 ? 8 : 1);
    case 2048: return s < 2 ? 16 : (s == 2 ? 8 : 1);
    case 4096: return s < 3 ? 16 : 1;
    // mixed radix: odd factors first (power-of-two strides for the remaining passes), the last pass
    // contiguous with an even radix
    case 96:   return s == 0 ? 3 : (s == 1 ? 8 : (s == 2 ? 4 : 1));
    case 160:  return s == 0 ? 5 : (s == 1 ? 8 : (s == 2 ? 4 : 1));
    case 192:  return s == 0 ? 3 : (s < 3 ? 8 : 1);
    case 320:  return s == 0 ? 5 : (s < 3 ? 8 : 1);
    case 384:  return s == 0 ? 3 : (s == 1 ? 16 : (s == 2 ? 8 : 1));
    case 480:  return s == 0 ? 3 : (s == 1 ? 5 : (s == 2 ? 8 : (s == 3 ? 4 : 1)));
    case 640:  return s == 0 ? 5 : (s == 1 ? 16 : (s == 2 ? 8 : 1));
    case 768:  return s == 0 ? 3 : (s < 3 ? 16 : 1);
    case 960:  return s == 0 ? 3 : (s == 1 ? 5 : (s < 4 ? 8 : 1));
    case 1280: return s == 0 ? 5 : (s < 3 ? 16 : 1);
    case 1536: return s == 0 ? 3 : (s < 4 ? 8 : 1);
    case 1920: return s == 0 ? 3 : (s == 1 ? 5 : (s == 2 ? 16 : (s == 3 ? 8 : 1)));
    default:   return 1;
  }
}
ADMMTV_HD constexpr int plan_stages(int L, int P = 0) {
  int n = 0;
  while (n < 4 && plan_radix(L, n, P) > 1) ++n;
  return n;
}
// length of the sub-transforms stage s works on: L / prod_{q<s} R_q
ADMMTV_HD constexpr int plan_sublen(int L, int s, int P = 0) {
  int len = L;
  for (int q = 0; q < s; ++q) len /= plan_radix(L, q, P);
  return len;
}
ADMMTV_HD constexpr bool plan_supported(int L) { return plan_stages(L) >= 2; }

// Supported transform lengths are addressed by a small id so that kernels stay templated on one int:
// ids 5..12 are the powers of two 32..4096, ids 20..31 the 3- and 5-smooth lengths below.
ADMMTV_HD constexpr int dim_len(int id) {
  switch (id) {
    case 20: return 96;   case 21: return 160;  case 22: return 192;  case 23: return 320;
    case 24: return 384;  case 25: return 480;  case 26: return 640;  case 27: return 768;
    case 28: return 960;  case 29: return 1280; case 30: return 1536; case 31: return 1920;
    default: return (id >= 5 && id <= 12) ? (1 << id) : 0;
  }
}
ADMMTV_HD constexpr bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
// id 0 = no register plan: the generic kernels (generic_kernels.cuh) handle any length 1..kMaxGenericLen
constexpr int kMaxGenericLen = 4096;
ADMMTV_HD inline int dim_id(int L) {
  for (int id = 5; id <= 12; ++id)
    if (dim_len(id) == L) return id;
  for (int id = 20; id <= 31; ++id)
    if (dim_len(id) == L) return id;
  return (L >= 1 && L <= kMaxGenericLen) ? 0 : -1;
}

// frequency index held at storage position p after the forward (DIF) passes
// (`planned` = false: the generic path keeps spectra in natural order)
ADMMTV_HD inline int pos_to_freq(int L, int p, bool planned = true, int P = 0) {
  if (!planned || !plan_supported(L)) return p;
  int k = 0, mult = 1, len = L;
  for (int s = 0; s < 4; ++s) {
    int R = plan_radix(L, s, P);
    if (R <= 1) break;
    len /= R;
    int m = p / len;
    p -= m * len;
    k += m * mult;
    mult *= R;
  }
  return k;
}

// storage position of frequency index k (inverse of pos_to_freq for a planned length)
ADMMTV_HD inline int freq_to_pos(int L, int k, int P = 0) {
  if (!plan_supported(L)) return k;
  int p = 0, len = L;
  for (int s = 0; s < 4; ++s) {
    int R = plan_radix(L, s, P);
    if (R <= 1) break;
    len /= R;
    p += (k % R) * len;
    k /= R;
  }
  return p;
}

// ------------------------------------------------------------------------------------------
// complex helpers (float2 = re, im)
// ------------------------------------------------------------------------------------------
// Blackwell has packed dual-fp32 arithmetic (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2): one issue slot for both
// components of a complex add.  The butterflies are more than half additions and the iteration kernels are issue-slot-bound as much
// as DRAM-bound, so complex add / subtract use it; results are bit-identical to the scalar form (same
// round-to-nearest adds).  The mov.b64 packs below are register-pair renamings, not instructions.
#ifndef ADMMTV_F32X2
#define ADMMTV_F32X2 1
#endif
#if ADMMTV_F32X2 && !defined(ADMMTV_EMU)
ADMMTV_DI float2 cadd(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  float2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
ADMMTV_DI float2 csub(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  float2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
#else
ADMMTV_DI float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
ADMMTV_DI float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
ADMMTV_DI float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
ADMMTV_DI float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
ADMMTV_DI float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// multiply by exp(-+ 2*pi*i * E/16)  (forward: minus sign; INV: plus sign), E compile time
template <int E, bool INV>
ADMMTV_DI float2 mul_w16(float2 a) {
  constexpr float C1 = 0.92387953251128674f;  // cos(pi/8)
  constexpr float S1 = 0.38268343236508977f;  // sin(pi/8)
  constexpr float H = 0.70710678118654752f;   // sqrt(1/2)
  constexpr int e = E & 15;
  if constexpr (e == 0) return a;
  else if constexpr (e == 4) return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
  else if constexpr (e == 8) return make_float2(-a.x, -a.y);
  else if constexpr (e == 12) return INV ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
  else if constexpr (e == 2)
    return INV ? make_float2((a.x - a.y) * H, (a.x + a.y) * H) : make_float2((a.x + a.y) * H, (a.y - a.x) * H);
  else if constexpr (e == 6)
    return INV ? make_float2(-(a.x + a.y) * H, (a.x - a.y) * H) : make_float2((a.y - a.x) * H, -(a.x + a.y) * H);
  else {
    // general: w = (c, -+s)
    constexpr float c = (e == 1) ? C1 : (e == 3) ? S1 : (e == 5) ? -S1 : (e == 7) ? -C1 : 0.f;
    constexpr float s = (e == 1) ? S1 : (e == 3) ? C1 : (e == 5) ? C1 : (e == 7) ? S1 : 0.f;
    static_assert(e == 1 || e == 3 || e == 5 || e == 7, "mul_w16: unsupported exponent");
    // forward w = c - i s ; inverse w = c + i s
    return INV ? make_float2(a.x * c - a.y * s, a.x * s + a.y * c) : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
  }
}

// multiply by exp(-+ 2*pi*i * E/32): even exponents are the W16 cases above (identical code, identical bits)
template <int E, bool INV>
ADMMTV_DI float2 mul_w32(float2 a) {
  constexpr int e = E & 31;
  if constexpr ((e & 1) == 0) return mul_w16<e / 2, INV>(a);
  else {
    constexpr float C1 = 0.98078528040323043f, S1 = 0.19509032201612825f;   // cos, sin (pi/16)
    constexpr float C3 = 0.83146961230254524f, S3 = 0.55557023301960218f;   // cos, sin (3 pi/16)
    // angle = e * pi/16, e odd in 1..15 (the DIF butterflies only use the upper half-plane exponents)
    static_assert(e < 16, "mul_w32: exponent out of the butterfly range");
    constexpr float c = e == 1 ? C1 : e == 3 ? C3 : e == 5 ? S3 : e == 7 ? S1 : e == 9 ? -S1 : e == 11 ? -S3 : e == 13 ? -C3 : -C1;
    constexpr float s = e == 1 ? S1 : e == 3 ? S3 : e == 5 ? C3 : e == 7 ? C1 : e == 9 ? C1 : e == 11 ? C3 : e == 13 ? S3 : S1;
    return INV ? make_float2(a.x * c - a.y * s, a.x * s + a.y * c) : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
  }
}

// ------------------------------------------------------------------------------------------
// size-R DFT in registers, natural order in -> natural order out (recursive radix-2 DIF)
// ------------------------------------------------------------------------------------------
template <int R, bool INV>
struct Dft {
  static ADMMTV_DI void run(float2* a) {
    float2 s[R / 2], d[R / 2];
    unroll_half<0>(a, s, d);
    Dft<R / 2, INV>::run(s);
    Dft<R / 2, INV>::run(d);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      a[2 * k] = s[k];
      a[2 * k + 1] = d[k];
    }
  }
  template <int T>
  static ADMMTV_DI void unroll_half(const float2* a, float2* s, float2* d) {
    if constexpr (T < R / 2) {
      s[T] = cadd(a[T], a[T + R / 2]);
      d[T] = mul_w32<T * (32 / R), INV>(csub(a[T], a[T + R / 2]));
      unroll_half<T + 1>(a, s, d);
    }
  }
};
template <bool INV>
struct Dft<1, INV> {
  static ADMMTV_DI void run(float2*) {}
};
// radix 3:  X1,2 = a0 - (a1+a2)/2 -+ i (sqrt3/2)(a1-a2)   (forward; the inverse swaps them)
template <bool INV>
struct Dft<3, INV> {
  static ADMMTV_DI void run(float2* a) {
    constexpr float S = 0.86602540378443865f;
    const float2 t1 = cadd(a[1], a[2]);
    const float2 t2 = make_float2(a[0].x - 0.5f * t1.x, a[0].y - 0.5f * t1.y);
    const float2 d = csub(a[1], a[2]);
    const float2 r = make_float2(S * d.y, -S * d.x);  // -i S d
    a[0] = cadd(a[0], t1);
    a[1] = INV ? csub(t2, r) : cadd(t2, r);
    a[2] = INV ? cadd(t2, r) : csub(t2, r);
  }
};
// radix 5
template <bool INV>
struct Dft<5, INV> {
  static ADMMTV_DI void run(float2* a) {
    constexpr float C1 = 0.30901699437494742f, C2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
    constexpr float S1 = 0.95105651629515357f, S2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
    const float2 t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
    const float2 m1 = make_float2(a[0].x + C1 * t1.x + C2 * t2.x, a[0].y + C1 * t1.y + C2 * t2.y);
    const float2 m2 = make_float2(a[0].x + C2 * t1.x + C1 * t2.x, a[0].y + C2 * t1.y + C1 * t2.y);
    const float2 n1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
    const float2 n2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
    const float2 r1 = make_float2(n1.y, -n1.x), r2 = make_float2(n2.y, -n2.x);  // -i n
    a[0] = cadd(a[0], cadd(t1, t2));
    a[1] = INV ? csub(m1, r1) : cadd(m1, r1);
    a[4] = INV ? cadd(m1, r1) : csub(m1, r1);
    a[2] = INV ? csub(m2, r2) : cadd(m2, r2);
    a[3] = INV ? cadd(m2, r2) : csub(m2, r2);
  }
};

// p[m] = w^m, m = 0..R-1, product tree of depth log2(m)
template <int R>
ADMMTV_DI void twiddle_powers(float2 w, float2* p) {
  p[0] = make_float2(1.f, 0.f);
  p[1] = w;
#pragma unroll
  for (int m = 2; m < R; ++m) p[m] = cmul(p[m / 2], p[m - m / 2]);
}

// One stage's work item: which line elements it touches and its twiddle base.
//   elements: base + m*stride, m = 0..R-1 ; twiddle W_LS^(t*m) = tw[t*(L/LS)]^m
template <int L, int S, int P = 0>
struct Stage {
  static constexpr int R = plan_radix(L, S, P);
  static constexpr int LS = plan_sublen(L, S, P);
  static constexpr int STRIDE = LS / R;  // also the number of distinct t
  static constexpr int ITEMS = L / R;    // work items per line
  static constexpr bool HAS_TW = STRIDE > 1;
  static ADMMTV_DI int base(int wi) { return (wi / STRIDE) * LS + (wi % STRIDE); }
  static ADMMTV_DI int tindex(int wi) { return (wi % STRIDE) * (L / LS); }
};

// forward pass on registers: a <- twiddle( DFT_R(a) )
template <int L, int S, int P = 0>
ADMMTV_DI void stage_fwd(float2* a, const float2* p /*powers, only if HAS_TW*/) {
  using St = Stage<L, S, P>;
  Dft<St::R, false>::run(a);
  if constexpr (St::HAS_TW) {
#pragma unroll
    for (int m = 1; m < St::R; ++m) a[m] = cmul(a[m], p[m]);
  }
}
// inverse pass on registers: a <- IDFT_R( conj-twiddle(a) )   (p holds conj powers)
template <int L, int S, int P = 0>
ADMMTV_DI void stage_inv(float2* a, const float2* p) {
  using St = Stage<L, S, P>;
  if constexpr (St::HAS_TW) {
#pragma unroll
    for (int m = 1; m < St::R; ++m) a[m] = cmul(a[m], p[m]);
  }
  Dft<St::R, true>::run(a);
}

// twiddle powers for work item wi of stage S (forward sign, or conjugated for the inverse)
template <int L, int S, bool INV, int P = 0>
ADMMTV_DI void stage_twiddles(int wi, const float2* __restrict__ tw, float2* p) {
  using St = Stage<L, S, P>;
  if constexpr (St::HAS_TW) {
    float2 w = tw[St::tindex(wi)];
    if (INV) w.y = -w.y;
    twiddle_powers<St::R>(w, p);
  }
}

}  // namespace admmtv
