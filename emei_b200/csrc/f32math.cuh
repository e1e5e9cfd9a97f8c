// Lean float32 math for the tolerance-mode (f32) kernels.
//
// The f32 kernels are bounded by warp-instruction issue, not by HBM (ncu, profiles/r01_*): CUDA's
// sincosf + four IEEE divides cost ~135 SASS instructions per Euler sub-step.  These replacements
// cost ~40 and stay well inside the 1e-5 rel + 1e-6 abs per-step envelope of BASELINE.json
// (measured by tools/f32_study, which compiles this very header for the host):
//   * sincos: Cody-Waite 3-constant reduction to [-pi/4, pi/4] + degree-7/8 minimax polynomials
//     (the classic single-precision kernels), quadrant fix-up with integer ops; |x| > 1e5 or
//     non-finite x falls back to sincosf (Payne-Hanek) on a cold branch.
//   * divisions by constants become multiplications by the double-rounded reciprocal; the one
//     data-dependent division is one MUFU.RCP (<= 1 ulp) and a multiply.
// Everything is __host__ __device__ so the numerics study runs the identical arithmetic on the CPU
// (fmaf is exact on both sides; only MUFU.RCP is replaced by 1.0f/x on the host).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace emei {
namespace f32 {

#ifndef EMEI_F32_STUDY_RCP_ULP
#define EMEI_F32_STUDY_RCP_ULP 0
#endif
#if defined(__CUDA_ARCH__)
#define EMEI_F32_DEVICE 1
#else
#define EMEI_F32_DEVICE 0
#endif

__host__ __device__ __forceinline__ uint32_t f2u(float f) {
#if EMEI_F32_DEVICE
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
__host__ __device__ __forceinline__ float u2f(uint32_t u) {
#if EMEI_F32_DEVICE
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// 1/x for normal x of moderate magnitude: one MUFU.RCP (documented max error 1 ulp).  A Newton
// step would halve that error; it is not worth 2 of the ~40 instructions of a sub-step (the
// float32 rounding of the state dominates the per-step error budget: tools/f32_study).
__host__ __device__ __forceinline__ float rcp_fast(float x) {
#if EMEI_F32_DEVICE
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  // host study: worst case of the device instruction = the correctly rounded reciprocal off by 1 ulp
  const float r = 1.0f / x;
  return EMEI_F32_STUDY_RCP_ULP == 0 ? r : u2f(f2u(r) + EMEI_F32_STUDY_RCP_ULP);
#endif
}

constexpr float kTwoOverPi = 0.636619772367581343f;
// pi/2 = hi + mid + lo, each the float32 nearest to the running remainder.  The FMA forms j*hi
// exactly, and x - j*hi is exactly representable for |x| <= 1e5 (multiple of 2^-23, magnitude < 1).
constexpr float kPio2Hi = 1.57079637050628662109e+00f;
constexpr float kPio2Mid = -4.37113882867379288655e-08f;
constexpr float kPio2Lo = -1.71512451000588187280e-15f;
constexpr float kRoundMagic = 12582912.0f;  // 1.5 * 2^23: adding it rounds to nearest integer
constexpr float kSinCosFastMax = 100000.0f;
constexpr float kSinCosSaneMax = 4.0e6f;  // |x * 2/pi| < 2^22: the magic-number rounding is still exact

// minimax coefficients on [-pi/4, pi/4] (single-precision sin/cos kernels)
constexpr float kS0 = -1.6666654611e-1f, kS1 = 8.3321608736e-3f, kS2 = -1.9515295891e-4f;
constexpr float kC0 = 4.166664568298827e-2f, kC1 = -1.388731625493765e-3f, kC2 = 2.443315711809948e-5f;

struct Reduced {
  float r;     // x - j*pi/2, |r| <= pi/4 (+ rounding)
  uint32_t q;  // j mod 4 (two's complement low bits)
};

__host__ __device__ __forceinline__ Reduced reduce_pio2(float x) {
  const float t = fmaf(x, kTwoOverPi, kRoundMagic);
  const float j = t - kRoundMagic;
  float r = fmaf(-j, kPio2Hi, x);
  r = fmaf(-j, kPio2Mid, r);
  r = fmaf(-j, kPio2Lo, r);
  return {r, f2u(t)};
}

__host__ __device__ __forceinline__ float sin_poly(float r, float r2) {
  float p = fmaf(kS2, r2, kS1);
  p = fmaf(p, r2, kS0);
  return fmaf(r * r2, p, r);
}
__host__ __device__ __forceinline__ float cos_poly(float r2) {
  float p = fmaf(kC2, r2, kC1);
  p = fmaf(p, r2, kC0);
  p = fmaf(p, r2, -0.5f);
  return fmaf(p, r2, 1.0f);
}

// sin and cos of x WITHOUT a range guard: valid (and ~1 ulp) for |x| <= 1e5, sane (quadrant-correct,
// error growing to ~3e-6, far below the float32 spacing of x there) up to kSinCosSaneMax; NaN/Inf -> NaN.
// Callers guard once per env step (see cartpole_f32.cuh), not once per evaluation.
__host__ __device__ __forceinline__ void sincos_core(float x, float* s, float* c) {
  const Reduced m = reduce_pio2(x);
  const float r2 = m.r * m.r;
  const float ps = sin_poly(m.r, r2);
  const float pc = cos_poly(r2);
  const bool swap = (m.q & 1u) != 0;
  const float sv = swap ? pc : ps;
  const float cv = swap ? ps : pc;
  // quadrant signs: sin negative for q in {2,3}; cos negative for q in {1,2}
  const uint32_t ssign = (m.q & 2u) << 30;
  const uint32_t csign = ((m.q + 1u) & 2u) << 30;
  *s = u2f(f2u(sv) ^ ssign);
  *c = u2f(f2u(cv) ^ csign);
}
__host__ __device__ __forceinline__ float cos_core(float x) {
  const Reduced m = reduce_pio2(x);
  const float r2 = m.r * m.r;
  const float v = (m.q & 1u) ? sin_poly(m.r, r2) : cos_poly(r2);
  return u2f(f2u(v) ^ (((m.q + 1u) & 2u) << 30));
}

__host__ __device__ __forceinline__ void sincos_libm(float x, float* s, float* c) {
#if EMEI_F32_DEVICE
  sincosf(x, s, c);
#else
  *s = sinf(x);
  *c = cosf(x);
#endif
}

// guarded versions: fast path for |x| <= 1e5, libm-grade (Payne-Hanek) slow path otherwise
__host__ __device__ __forceinline__ void sincos_fast(float x, float* s, float* c) {
  if (!(fabsf(x) <= kSinCosFastMax)) {
    sincos_libm(x, s, c);
    return;
  }
  sincos_core(x, s, c);
}
__host__ __device__ __forceinline__ float cos_fast(float x) {
  if (!(fabsf(x) <= kSinCosFastMax)) return cosf(x);
  return cos_core(x);
}

// ---------------------------------------------------------------------------------------------
// cart-pole acceleration, algebraically identical to cartpole.py:51-58 with the constant
// divisions folded into pre-rounded reciprocals:
//   temp      = force/mt + (pml/mt) * w^2 * s
//   theta_acc = (g*s - c*temp) / (l*4/3 - (l*mp/mt) * c^2)
//   x_acc     = temp - (pml/mt) * theta_acc * c
// ---------------------------------------------------------------------------------------------
struct CartPoleK {
  float g, kpm /* pml/mt */, inv_mt, den0 /* l*4/3 */, den1 /* l*mp/mt */, dt;
};

template <bool LIBM>
__host__ __device__ __forceinline__ void cartpole_substep(float& x, float& xd, float& th, float& w, float f_mt,
                                                          float sgn, const CartPoleK& k) {
  float s, c;
  if (LIBM)
    sincos_libm(th, &s, &c);
  else
    sincos_core(th, &s, &c);
  s *= sgn;  // analytic inverted pendulum (swing-up models hang the pole down): theta_cartpole = theta + pi
  c *= sgn;
  const float temp = fmaf(k.kpm * (w * w), s, f_mt);
  const float num = fmaf(k.g, s, -(c * temp));
  const float den = fmaf(-k.den1, c * c, k.den0);
  const float th_acc = num * rcp_fast(den);
  const float x_acc = fmaf(-k.kpm, th_acc * c, temp);
  x = fmaf(xd, k.dt, x);
  xd = fmaf(x_acc, k.dt, xd);
  th = fmaf(w, k.dt, th);
  w = fmaf(th_acc, k.dt, w);
}

}  // namespace f32
}  // namespace emei
