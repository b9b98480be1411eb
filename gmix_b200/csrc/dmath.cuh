// Exact device arithmetic for the gmix per-bit path.
//
// The reference computes everything in source-order IEEE fp32 (strict build: no FMA contraction,
// SURVEY.md section 0.4 / appendix C) and calls glibc 2.39 libm for expf/logf/tanhf
// (reference src/mixer/sigmoid.cpp:5-13, src/models/lstm.cpp:115, src/models/lstm-layer.cpp:208,215).
// To make GPU-compressed streams byte-identical, this header provides
//   * f_add/f_sub/f_mul/f_div/f_sqrt: single IEEE-rounded fp32 operations that nvcc can never
//     contract into FFMA (round-to-nearest intrinsics), and
//   * gm_expf/gm_logf/gm_tanhf: bit-exact re-implementations of the glibc 2.39 algorithms
//     (expf/logf: the double-precision table algorithms of sysdeps/ieee754/flt-32/e_expf.c and
//     e_logf.c in their FMA ifunc variant, which is what an AVX2+FMA host runs; tanhf/expm1f:
//     the fdlibm single-precision code, s_tanhf.c / s_expm1f.c). SURVEY.md appendix E lists the
//     constants; tests/native/dmath_check.cpp sweeps all 2^32 inputs against the host libm and
//     gmx_selftest_math() repeats the sweep on the GPU.
// The same source compiles for the host (plain C++ with -ffp-contract=off) so the CPU-side
// sweep in tests/ exercises exactly this code.
#ifndef GMIX_B200_DMATH_CUH_
#define GMIX_B200_DMATH_CUH_

#include <stdint.h>
#include <string.h>
#if !defined(__CUDACC__)
#include <math.h>
#endif

#if defined(__CUDACC__)
#define GMX_HD __host__ __device__ __forceinline__
// The libm bodies are called from a dozen sites of the stream kernel; one out-of-line copy each keeps
// the per-bit path's code small (it is instruction-cache bound).
#define GMX_HD_OUTLINE static __host__ __device__ __noinline__
#else
#define GMX_HD inline
#define GMX_HD_OUTLINE inline
#endif

namespace gmx {

GMX_HD float f_add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
GMX_HD float f_sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
GMX_HD float f_mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
GMX_HD float f_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
GMX_HD float f_sqrt(float a) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
GMX_HD double d_mul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
GMX_HD double d_add(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
GMX_HD double d_sub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
GMX_HD double d_div(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
GMX_HD double d_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}

GMX_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
GMX_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
GMX_HD uint64_t d2u(double d) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
GMX_HD double u2d(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}

// 2^(i/32) table of glibc's __exp2f_data (N = 32): T[i] = bits(2^(i/32)) - (i << 47).
#define GMX_EXP2T_VALUES \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull
// {invc, logc} pairs of glibc's __logf_data (16 entries).
#define GMX_LOGT_VALUES \
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, \
    0x1.49539f0f010bp+0,  -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, \
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0,  -0x1.1aa2bc79c81p-3,   \
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, \
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0,               0x0p+0,                \
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aap-1,  0x1.c5e53aa362eb4p-4,  \
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3,   \
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2

namespace tables {
static const uint64_t kExp2T_host[32] = {GMX_EXP2T_VALUES};
static const double kLogT_host[32] = {GMX_LOGT_VALUES};
#if defined(__CUDACC__)
static __device__ const uint64_t kExp2T_dev[32] = {GMX_EXP2T_VALUES};
static __device__ const double kLogT_dev[32] = {GMX_LOGT_VALUES};
#endif
GMX_HD uint64_t Exp2T(int i) {
#ifdef __CUDA_ARCH__
  return kExp2T_dev[i];
#else
  return kExp2T_host[i];
#endif
}
GMX_HD double LogT(int i) {
#ifdef __CUDA_ARCH__
  return kLogT_dev[i];
#else
  return kLogT_host[i];
#endif
}
}  // namespace tables

// glibc 2.39 expf, FMA variant.
GMX_HD_OUTLINE float gm_expf(float x) {
  const uint32_t ux = f2u(x);
  const uint32_t abstop = (ux >> 20) & 0x7ff;
  if (abstop >= 0x42b) {  // |x| >= 88 or non-finite
    if (ux == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8) return f_add(x, x);
    if (x > 0x1.62e42ep6f) return u2f(0x7f800000u);  // overflow -> +inf
    if (x < -0x1.9fe368p6f) return 0.0f;             // underflow -> 0
    if (x < -0x1.9d1d9ep6f) return u2f(1u);          // 0x1p-149f
  }
  const double InvLn2N = 0x1.71547652b82fep+5, Shift = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;
  const double xd = (double)x;
  double kd = d_fma(InvLn2N, xd, Shift);
  const uint64_t ki = d2u(kd);
  kd = d_sub(kd, Shift);
  const double r = d_fma(InvLn2N, xd, -kd);
  const double s = u2d(tables::Exp2T((int)(ki & 31)) + (ki << 47));
  const double z = d_fma(C0, r, C1);
  const double r2 = d_mul(r, r);
  double y = d_fma(C2, r, 1.0);
  y = d_fma(z, r2, y);
  y = d_mul(y, s);
  return (float)y;
}

// glibc 2.39 logf, FMA variant, positive normal arguments only: every
// call site clamps its argument first (Sigmoid::Logit, sigmoid.cpp:7-12 => x in [1e-4, 1e4]).
GMX_HD_OUTLINE float gm_logf(float x) {
  const uint32_t ix = f2u(x);
  if (ix == 0x3f800000u) return 0.0f;
  const double Ln2 = 0x1.62e42fefa39efp-1;
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (tmp >> 19) & 15;
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double invc = tables::LogT(2 * i), logc = tables::LogT(2 * i + 1);
  const double z = (double)u2f(iz);
  const double r = d_fma(z, invc, -1.0);
  const double y0 = d_fma((double)k, Ln2, logc);
  const double r2 = d_mul(r, r);
  double y = d_fma(A1, r, A2);
  y = d_fma(A0, r2, y);
  y = d_fma(y, r2, d_add(y0, r));
  return (float)y;
}

// fdlibm expm1f as shipped in glibc (s_expm1f.c): all operations single precision, no FMA.
GMX_HD float gm_expm1f(float x) {
  const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f;
  const float ln2_hi = u2f(0x3f317180u), ln2_lo = u2f(0x3717f7d1u), invln2 = u2f(0x3fb8aa3bu);
  const float Q1 = u2f(0xbd088889u), Q2 = u2f(0x3ad00d01u), Q3 = u2f(0xb8a670cdu);
  const float Q4 = u2f(0x36867e54u), Q5 = u2f(0xb457edbbu);
  const float o_threshold = u2f(0x42b17180u);
  float y, hi, lo, c = 0.0f, t, e, hxs, hfx, r1;
  int32_t k;
  uint32_t hx = f2u(x);
  const uint32_t xsb = hx & 0x80000000u;
  hx &= 0x7fffffffu;
  if (hx >= 0x4195b844u) {  // |x| >= 27 ln2
    if (hx >= 0x42b17218u) {  // |x| >= 88.72
      if (hx > 0x7f800000u) return f_add(x, x);
      if (hx == 0x7f800000u) return xsb == 0 ? x : -1.0f;
      if (x > o_threshold) return f_mul(huge, huge);
    }
    if (xsb != 0) {
      if (f_add(x, tiny) < 0.0f) return f_sub(tiny, one);
    }
  }
  if (hx > 0x3eb17218u) {  // |x| > 0.5 ln2
    if (hx < 0x3F851592u) {  // |x| < 1.5 ln2
      if (xsb == 0) { hi = f_sub(x, ln2_hi); lo = ln2_lo; k = 1; }
      else { hi = f_add(x, ln2_hi); lo = -ln2_lo; k = -1; }
    } else {
      k = (int32_t)f_add(f_mul(invln2, x), xsb == 0 ? 0.5f : -0.5f);
      t = (float)k;
      hi = f_sub(x, f_mul(t, ln2_hi));
      lo = f_mul(t, ln2_lo);
    }
    x = f_sub(hi, lo);
    c = f_sub(f_sub(hi, x), lo);
  } else if (hx < 0x33000000u) {  // |x| < 2^-25
    t = f_add(huge, x);
    return f_sub(x, f_sub(t, f_add(huge, x)));
  } else {
    k = 0;
  }
  hfx = f_mul(0.5f, x);
  hxs = f_mul(x, hfx);
  r1 = f_add(one, f_mul(hxs, f_add(Q1, f_mul(hxs, f_add(Q2, f_mul(hxs, f_add(Q3, f_mul(hxs, f_add(Q4, f_mul(hxs, Q5)))))))))
  );
  t = f_sub(3.0f, f_mul(r1, hfx));
  e = f_mul(hxs, f_div(f_sub(r1, t), f_sub(6.0f, f_mul(x, t))));
  if (k == 0) return f_sub(x, f_sub(f_mul(x, e), hxs));
  e = f_sub(f_mul(x, f_sub(e, c)), c);
  e = f_sub(e, hxs);
  if (k == -1) return f_sub(f_mul(0.5f, f_sub(x, e)), 0.5f);
  if (k == 1) {
    if (x < -0.25f) return f_mul(-2.0f, f_sub(e, f_add(x, 0.5f)));
    return f_add(one, f_mul(2.0f, f_sub(x, e)));
  }
  if (k <= -2 || k > 56) {
    y = f_sub(one, f_sub(e, x));
    y = u2f(f2u(y) + ((uint32_t)k << 23));
    return f_sub(y, one);
  }
  if (k < 23) {
    t = u2f(0x3f800000u - (0x1000000u >> k));  // 1 - 2^-k
    y = f_sub(t, f_sub(e, x));
    y = u2f(f2u(y) + ((uint32_t)k << 23));
  } else {
    t = u2f((uint32_t)(0x7f - k) << 23);  // 2^-k
    y = f_sub(x, f_add(e, t));
    y = f_add(y, one);
    y = u2f(f2u(y) + ((uint32_t)k << 23));
  }
  return y;
}

// fdlibm tanhf as shipped in glibc (s_tanhf.c).
GMX_HD_OUTLINE float gm_tanhf(float x) {
  const float one = 1.0f, two = 2.0f, tiny = 1.0e-30f;
  const uint32_t jx = f2u(x);
  const uint32_t ix = jx & 0x7fffffffu;
  float z;
  if (ix >= 0x7f800000u) {
    if ((int32_t)jx >= 0) return f_add(f_div(one, x), one);
    return f_sub(f_div(one, x), one);
  }
  if (ix < 0x41b00000u) {  // |x| < 22
    if (ix == 0) return x;
    if (ix < 0x24000000u) return f_mul(x, f_add(one, x));  // |x| < 2^-55
    const float ax = u2f(ix);
    if (ix >= 0x3f800000u) {
      const float t = gm_expm1f(f_mul(two, ax));
      z = f_sub(one, f_div(two, f_add(t, two)));
    } else {
      const float t = gm_expm1f(f_mul(-two, ax));
      z = f_div(-t, f_add(t, two));
    }
  } else {
    z = f_sub(one, tiny);
  }
  return (int32_t)jx >= 0 ? z : -z;
}

// reference src/mixer/sigmoid.cpp:5 — 1 / (1 + exp(-p)), all fp32.
GMX_HD float Logistic(float p) { return f_div(1.0f, f_add(1.0f, gm_expf(-p))); }
// reference src/mixer/sigmoid.cpp:7-13 — comparisons against the double literals, assignment
// of the literal rounded to float.
GMX_HD float Logit(float p) {
  if ((double)p < 0.0001) p = 0.0001f;
  else if ((double)p > 0.9999) p = 0.9999f;
  return gm_logf(f_div(p, f_sub(1.0f, p)));
}

}  // namespace gmx
#endif  // GMIX_B200_DMATH_CUH_
