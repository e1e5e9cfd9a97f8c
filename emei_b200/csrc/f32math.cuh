// Lean float32 math for the tolerance-mode (f32) kernels.
//
// CUDA's sincosf + four IEEE divides cost ~135 SASS instructions per Euler sub-step, which made the f32 step
// issue-bound.  These replacements cost ~40 (and a sub-step after the first ~30, see cartpole_integrate) and stay
// well inside the 1e-5 rel + 1e-6 abs per-step envelope of BASELINE.json
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

// ---------------------------------------------------------------------------------------------
// One source for two instruction widths.  The arithmetic below is written over a value type V:
//   V = float : scalar FFMA / FMUL / FADD (host and device; rollout kernels, charged ball, the host study)
//   V = f2    : two envs per thread in one 64-bit register pair, Blackwell's packed fma/mul/add.rn.f32x2
//               (SASS FFMA2 / FMUL2 / FADD2): one issue slot does the FP work of two envs.  The step kernel
//               is bound by warp-instruction issue (ncu: issue 75 %, fma pipe 43 %, alu pipe 36 %), and
//               FFMA2 occupies the fma pipe for two cycles but the scheduler for one, so packing moves the
//               bound from "issue slots" (~250 / env-step) to "fma-pipe cycles" (~135 / env-step).
// Every f32x2 lane is an IEEE fma.rn / mul.rn / add.rn, so V = f2 produces exactly the bits of V = float:
// negations are exact (negated constants, negated intermediates and vneg() commute with round-to-nearest; ptxas
// folds a packed vneg into the consuming FFMA2's operand modifier).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
#if EMEI_F32_DEVICE
__host__ __device__ __forceinline__ float vmul(float a, float b) { return __fmul_rn(a, b); }
__host__ __device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }
__host__ __device__ __forceinline__ float vneg(float a) { return -a; }
#else
__host__ __device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__host__ __device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__host__ __device__ __forceinline__ float vneg(float a) { return -a; }
#endif
template <class V>
struct Splat;
template <>
struct Splat<float> {
  __host__ __device__ __forceinline__ static float of(float a) { return a; }
};
#pragma nv_exec_check_disable
template <class V>
__host__ __device__ __forceinline__ V vsplat(float a) {
  return Splat<V>::of(a);
}

#ifdef __CUDACC__
struct f2 {
  unsigned long long v;  // {lo = env A, hi = env B}
};
__device__ __forceinline__ f2 f2_pack(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 vfma(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return d;
}
__device__ __forceinline__ f2 vmul(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ f2 vadd(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
// exact negation of both lanes; ptxas folds it into the operand modifier of the consuming FFMA2 (-R.F32x2).
// (A packed subtraction is deliberately NOT offered: ptxas contracts mul.rn.f32x2 + sub.rn.f32x2 into one FFMA2,
// which would differ from the scalar __fmul_rn / __fsub_rn form by one rounding.)
__device__ __forceinline__ f2 vneg(f2 a) {
  float lo, hi;
  f2_unpack(a, lo, hi);
  return f2_pack(-lo, -hi);
}
template <>
struct Splat<f2> {
  __device__ __forceinline__ static f2 of(float a) { return f2_pack(a, a); }
};
#endif

// quadrant fix-up of one lane: (ps, pc) = (sin r, cos r) on the reduced argument, tbits = the bits of the
// magic-number sum (low bits = j mod 4); flip = 0 or 0x80000000 negates both results (the analytic inverted
// pendulum's swing-up models hang the pole down: theta_cartpole = theta + pi).
__host__ __device__ __forceinline__ void quadrant_fix(float ps, float pc, uint32_t tbits, uint32_t flip, float* s, float* c) {
  const bool swap = (tbits & 1u) != 0;
  const float sv = swap ? pc : ps;
  const float cv = swap ? ps : pc;
  // quadrant signs: sin negative for q in {2,3}; cos negative for q in {1,2}
  const uint32_t ssign = ((tbits & 2u) << 30) ^ flip;
  const uint32_t csign = (((tbits + 1u) & 2u) << 30) ^ flip;
  *s = u2f(f2u(sv) ^ ssign);
  *c = u2f(f2u(cv) ^ csign);
}

// reduced argument r = x - j*pi/2 (|r| <= pi/4 + rounding) and t = x*2/pi + magic (its low bits hold j mod 4)
#pragma nv_exec_check_disable
template <class V>
__host__ __device__ __forceinline__ void reduce_pio2(V x, V* r_out, V* t_out) {
  const V t = vfma(x, vsplat<V>(kTwoOverPi), vsplat<V>(kRoundMagic));
  const V j = vadd(t, vsplat<V>(-kRoundMagic));
  V r = vfma(j, vsplat<V>(-kPio2Hi), x);
  r = vfma(j, vsplat<V>(-kPio2Mid), r);
  r = vfma(j, vsplat<V>(-kPio2Lo), r);
  *r_out = r;
  *t_out = t;
}
#pragma nv_exec_check_disable
template <class V>
__host__ __device__ __forceinline__ V sin_poly(V r, V r2) {
  V p = vfma(vsplat<V>(kS2), r2, vsplat<V>(kS1));
  p = vfma(p, r2, vsplat<V>(kS0));
  return vfma(vmul(r, r2), p, r);
}
#pragma nv_exec_check_disable
template <class V>
__host__ __device__ __forceinline__ V cos_poly(V r2) {
  V p = vfma(vsplat<V>(kC2), r2, vsplat<V>(kC1));
  p = vfma(p, r2, vsplat<V>(kC0));
  p = vfma(p, r2, vsplat<V>(-0.5f));
  return vfma(p, r2, vsplat<V>(1.0f));
}

// sin and cos of x WITHOUT a range guard: valid (and ~1 ulp) for |x| <= 1e5, sane (quadrant-correct,
// error growing to ~3e-6, far below the float32 spacing of x there) up to kSinCosSaneMax; NaN/Inf -> NaN.
// Callers guard once per env step (see cartpole_f32.cuh), not once per evaluation.
__host__ __device__ __forceinline__ void sincos_core(float x, float* s, float* c, uint32_t flip = 0u) {
  float r, t;
  reduce_pio2<float>(x, &r, &t);
  const float r2 = vmul(r, r);
  quadrant_fix(sin_poly<float>(r, r2), cos_poly<float>(r2), f2u(t), flip, s, c);
}
__host__ __device__ __forceinline__ float cos_core(float x) {
  float r, t;
  reduce_pio2<float>(x, &r, &t);
  const float r2 = vmul(r, r);
  const uint32_t q = f2u(t);
  const float v = (q & 1u) ? sin_poly<float>(r, r2) : cos_poly<float>(r2);
  return u2f(f2u(v) ^ (((q + 1u) & 2u) << 30));
}
#ifdef __CUDACC__
__device__ __forceinline__ void sincos_core(f2 x, f2* s, f2* c, uint32_t flip = 0u) {
  f2 r, t;
  reduce_pio2<f2>(x, &r, &t);
  const f2 r2 = vmul(r, r);
  const f2 ps = sin_poly<f2>(r, r2), pc = cos_poly<f2>(r2);
  float psa, psb, pca, pcb, ta, tb, sa, sb, ca, cb;
  f2_unpack(ps, psa, psb);
  f2_unpack(pc, pca, pcb);
  f2_unpack(t, ta, tb);
  quadrant_fix(psa, pca, f2u(ta), flip, &sa, &ca);
  quadrant_fix(psb, pcb, f2u(tb), flip, &sb, &cb);
  *s = f2_pack(sa, sb);
  *c = f2_pack(ca, cb);
}
__device__ __forceinline__ f2 cos_core(f2 x) {  // both lanes evaluate both polynomials (packed), then select
  f2 r, t;
  reduce_pio2<f2>(x, &r, &t);
  const f2 r2 = vmul(r, r);
  const f2 ps = sin_poly<f2>(r, r2), pc = cos_poly<f2>(r2);
  float psa, psb, pca, pcb, ta, tb;
  f2_unpack(ps, psa, psb);
  f2_unpack(pc, pca, pcb);
  f2_unpack(t, ta, tb);
  const uint32_t qa = f2u(ta), qb = f2u(tb);
  const float va = (qa & 1u) ? psa : pca, vb = (qb & 1u) ? psb : pcb;
  return f2_pack(u2f(f2u(va) ^ (((qa + 1u) & 2u) << 30)), u2f(f2u(vb) ^ (((qb + 1u) & 2u) << 30)));
}
#endif

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

__host__ __device__ __forceinline__ float vrcp(float x) { return rcp_fast(x); }
#ifdef __CUDACC__
__device__ __forceinline__ f2 vrcp(f2 x) {
  float a, b;
  f2_unpack(x, a, b);
  return f2_pack(rcp_fast(a), rcp_fast(b));
}
#endif

// One forward-Euler sub-step given (s, c) = (sin, cos) of the CART-POLE angle.  nf_mt = -(force / m_total).
// Written with -temp and -x_acc so that the packed form needs no negation; bit for bit
//   temp = fma(kpm*(w*w), s, f_mt); num = fma(g, s, -(c*temp)); den = fma(-den1, c*c, den0);
//   th_acc = num * rcp(den); x_acc = fma(-kpm, th_acc*c, temp); x += xd*dt; xd += x_acc*dt; th += w*dt; w += th_acc*dt
#pragma nv_exec_check_disable
template <class V>
__host__ __device__ __forceinline__ void cartpole_euler(V& x, V& xd, V& th, V& w, V s, V c, V nf_mt, const CartPoleK& k) {
  const V ntemp = vfma(vmul(vsplat<V>(-k.kpm), vmul(w, w)), s, nf_mt);
  const V num = vfma(vsplat<V>(k.g), s, vmul(c, ntemp));
  const V den = vfma(vsplat<V>(-k.den1), vmul(c, c), vsplat<V>(k.den0));
  const V th_acc = vmul(num, vrcp(den));
  const V nx_acc = vfma(vsplat<V>(k.kpm), vmul(th_acc, c), ntemp);
  const V dt = vsplat<V>(k.dt);
  x = vfma(xd, dt, x);
  xd = vfma(nx_acc, vsplat<V>(-k.dt), xd);
  th = vfma(w, dt, th);
  w = vfma(th_acc, dt, w);
}

// ---------------------------------------------------------------------------------------------
// One env step = freq_rate sub-steps with ONE full sincos.  The reference evaluates sin/cos of the angle
// at every sub-step (cartpole.py:51-52); between sub-steps the angle moves by d = theta_dot * dt exactly
// (base_control.py:164), so (sin, cos) of the next angle follow from the angle-addition formulas with
// sin d, cos d from the [-pi/4, pi/4] kernels -- no range reduction and no quadrant fix-up (16 of the 47
// instructions of a packed sub-step were integer quadrant logic; ncu showed the step kernel bound by
// issue slots).  It also tracks the reference more closely for large angles: the reference's float64 angle
// is theta0 + sum(d) without the float32 rounding of the stored theta, and so is the rotated (sin, cos).
// The caller checks |theta0| <= kSinCosSaneMax and max |d| <= kDeltaMax (LaneMax) once per env step
// and redoes the rare env that fails with the libm path (cartpole_substep<true>).
// Returns cos of the FINAL cart-pole angle (the swing-up reward), flipped like the sub-step values.
// ---------------------------------------------------------------------------------------------
constexpr float kDeltaMax = 0.785398185253143310546875f;  // float32(pi/4): |theta_dot| <= 39 rad/s at dt = 0.02

template <class V>
struct LaneMax;
template <>
struct LaneMax<float> {
  float m = 0.0f;
  __host__ __device__ __forceinline__ void acc(float d) { m = fmaxf(m, fabsf(d)); }
};
#ifdef __CUDACC__
template <>
struct LaneMax<f2> {
  float a = 0.0f, b = 0.0f;
  __device__ __forceinline__ void acc(f2 d) {
    float x, y;
    f2_unpack(d, x, y);
    a = fmaxf(a, fabsf(x));
    b = fmaxf(b, fabsf(y));
  }
};
#endif

#pragma nv_exec_check_disable
template <class V, int FR>
__host__ __device__ __forceinline__ V cartpole_integrate(V& x, V& xd, V& th, V& w, V nf_mt, uint32_t flip, const CartPoleK& k,
                                                         int fr_runtime, LaneMax<V>& dmax) {
  V s, c;
  sincos_core(th, &s, &c, flip);
  const int fr = FR > 0 ? FR : fr_runtime;
#pragma unroll
  for (int sub = 0; sub < fr; ++sub) {
    const V d = vmul(w, vsplat<V>(k.dt));  // this sub-step's angle increment (th = fma(w, dt, th) in cartpole_euler)
    dmax.acc(d);
    cartpole_euler<V>(x, xd, th, w, s, c, nf_mt, k);
    const V d2 = vmul(d, d);
    const V sd = sin_poly<V>(d, d2), cd = cos_poly<V>(d2);
    const V s_next = vfma(s, cd, vmul(c, sd));
    c = vfma(c, cd, vneg(vmul(s, sd)));
    s = s_next;
  }
  return c;
}

// sub-step of the state (x, xd, th, w).  flip = 0x80000000 for the inverted pendulum's hanging models.
template <bool LIBM>
__host__ __device__ __forceinline__ void cartpole_substep(float& x, float& xd, float& th, float& w, float f_mt,
                                                          uint32_t flip, const CartPoleK& k) {
  float s, c;
  if (LIBM) {
    sincos_libm(th, &s, &c);
    s = u2f(f2u(s) ^ flip);
    c = u2f(f2u(c) ^ flip);
  } else {
    sincos_core(th, &s, &c, flip);
  }
  cartpole_euler<float>(x, xd, th, w, s, c, -f_mt, k);
}
#ifdef __CUDACC__
// packed sub-step with a full sincos: the previous generation, kept for the A/B variants of tools/kbench
__device__ __forceinline__ void cartpole_substep2(f2& x, f2& xd, f2& th, f2& w, f2 nf_mt, uint32_t flip, const CartPoleK& k) {
  f2 s, c;
  sincos_core(th, &s, &c, flip);
  cartpole_euler<f2>(x, xd, th, w, s, c, nf_mt, k);
}
#endif

}  // namespace f32
}  // namespace emei
