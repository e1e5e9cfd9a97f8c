// float32 (tolerance-mode) charged-ball step: freq_rate x update_state(_get_update_info(E))
// (charged_ball.py:54-82) with circle_to_free / free_to_circle / _get_angle / _angle_greater (:25-52),
// reward (:158-160) and terminal (:110-111), in place on the three state arrays.
//
// HBM-bound (56 B/env with uint8 actions): persistent grid (148 SMs x 8 CTAs x 256 threads), grid-stride,
// every load of an env issued before any math, 2048 resident threads/SM (<= 32 registers) so ~50 KB of
// loads are in flight per SM; reward partials stay in a register and are reduced ONCE per thread (the
// per-env warp-shuffle reduction of the first version made the kernel MIO-bound: ncu short-scoreboard
// stalls 54 %, 2.8 TB/s).  The common
// branches (on the circle: two sincos; free flight: four FMAs) use the lean float32 math of
// f32math.cuh; only a landing (rare) evaluates asin / fmod / sqrt through libm.
#pragma once
#include "common.cuh"
#include "f32math.cuh"

namespace emei {

struct ChargedBallF32Consts {
  float mg /* m*g */, inv_mr /* 1/(m*r) */, m_r /* m*r */, inv_m, g, r, inv_r, charge, h, land_thr, eps;
  int freq_rate;
};

inline ChargedBallF32Consts make_cb_f32_consts(const emei_charged_ball_params& p) {
  ChargedBallF32Consts k;
  k.mg = static_cast<float>(p.mass_ball * p.gravity_acc);
  k.inv_mr = static_cast<float>(1.0 / (p.mass_ball * p.radius));
  k.m_r = static_cast<float>(p.mass_ball * p.radius);
  k.inv_m = static_cast<float>(1.0 / p.mass_ball);
  k.g = static_cast<float>(p.gravity_acc);
  k.r = static_cast<float>(p.radius);
  k.inv_r = static_cast<float>(1.0 / p.radius);
  k.charge = static_cast<float>(p.charge);
  k.h = static_cast<float>(p.time_step / p.freq_rate);  // charged_ball.py:58,63
  k.land_thr = static_cast<float>(p.radius * p.radius + 0.001);  // :64
  k.eps = 1e-8f;
  k.freq_rate = p.freq_rate;
  return k;
}

// charged_ball.py:30-36 (cold path: only evaluated when a ball lands)
__device__ __noinline__ float cb_get_angle_f32(float x, float y, float r, float eps) {
  const float scale = sqrtf(x * x + y * y);
  const float a = asinf(x / (scale * r + eps));
  const float angle = (y > 0.f) ? a : (3.14159265358979323846f - a);
  return py_mod(angle, 6.28318530717958647692f);
}

template <int AK>
__global__ void __launch_bounds__(kBlock, 8)
    charged_ball_step_f32_kernel(uint8_t* on_circle, float2* circle, float4* free_state, const void* __restrict__ action,
                                 float* __restrict__ reward, uint8_t* __restrict__ done, double* stats, int64_t n,
                                 const ChargedBallF32Consts k) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kBlock;
  float r_acc = 0.f;  // <= a few hundred rewards in [-inf, 1] per thread; widened to double for the block reduction
  pdl_trigger();
  pdl_wait();
#pragma unroll 1
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; i < n; i += stride) {
    // all loads first
    const uint8_t on_u8 = on_circle[i];
    const float2 c2 = circle[i];
    float4 f = free_state[i];  // x, y, vx, vy
    const float a = load_action_f32<AK>(action, i);
    bool on = on_u8 != 0;
    float theta = c2.x, omega = c2.y;
    float E;
    if constexpr (AK <= EMEI_ACTION_DISCRETE_I64)
      E = a == 1.0f ? k.charge : -k.charge;  // charged_ball.py:155-156
    else
      E = k.charge * a;  // :169-170
    const bool fast_trig = fabsf(theta) <= f32::kSinCosFastMax;  // NaN / huge angles -> libm
    for (int sub = 0; sub < k.freq_rate; ++sub) {
      if (on) {
        // _get_update_info :72-78 + update_state :56-61 + circle_to_free :25-28
        float s, c;
        if (fast_trig) f32::sincos_core(theta, &s, &c); else sincosf(theta, &s, &c);
        const float theta_acc = fmaf(s, k.mg, c * E) * k.inv_mr;
        const bool flag = fmaf(s, E, k.m_r * (omega * omega)) < c * k.mg;  // evaluated on the pre-update state
        theta = fmaf(omega, k.h, theta);
        omega = fmaf(theta_acc, k.h, omega);
        float sn, cn;
        if (fast_trig && fabsf(theta) <= f32::kSinCosSaneMax) f32::sincos_core(theta, &sn, &cn); else sincosf(theta, &sn, &cn);
        f.x = sn * k.r;
        f.y = cn * k.r;
        f.z = omega * f.y;
        f.w = -omega * f.x;
        if (flag) on = false;
      } else {
        // _get_update_info :79-82 + update_state :62-66 + free_to_circle :44-52
        const float acc_x = E * k.inv_m;
        const float nx = fmaf(f.z, k.h, f.x), ny = fmaf(f.w, k.h, f.y);
        f.z = fmaf(acc_x, k.h, f.z);
        f.w = fmaf(-k.g, k.h, f.w);
        f.x = nx;
        f.y = ny;
        if (fmaf(f.x, f.x, f.y * f.y) > k.land_thr) {
          on = true;
          theta = cb_get_angle_f32(f.x, f.y, k.r, k.eps);
          const float v_angle = cb_get_angle_f32(f.z, f.w, k.r, k.eps);
          const bool greater = (fabsf(v_angle - theta) < 3.14159265358979323846f) ? (v_angle > theta) : (v_angle < theta);  // :38-42
          const float speed = sqrtf(fmaf(f.z, f.z, f.w * f.w)) * k.inv_r;
          omega = greater ? speed : -speed;
        }
      }
    }
    on_circle[i] = on ? 1 : 0;
    circle[i] = make_float2(theta, omega);
    free_state[i] = f;
    const float rew = 1.0f - sqrtf(fmaf(f.x, f.x, f.y * f.y)) * k.inv_r;  // charged_ball.py:158-160
    reward[i] = rew;
    done[i] = 0;  // charged_ball.py:110-111
    r_acc += rew;
  }
  block_stats_accumulate_counts(stats, static_cast<double>(r_acc), 0u);
}

inline void charged_ball_step_f32_dispatch(uint8_t* on_circle, float* circle, float* free_state, const void* action,
                                           float* reward, uint8_t* done, double* stats, int64_t n,
                                           const emei_charged_ball_params& p, cudaStream_t s) {
  const ChargedBallF32Consts k = make_cb_f32_consts(p);
  const int grid = persistent_grid(n, kBlock, 8);
  float2* c2 = reinterpret_cast<float2*>(circle);
  float4* f4 = reinterpret_cast<float4*>(free_state);
  switch (p.action_kind) {
#define EMEI_AK(A)                                                                                                 \
  case A:                                                                                                          \
    launch_pdl(charged_ball_step_f32_kernel<A>, grid, kBlock, s, on_circle, c2, f4, action, reward, done, stats, n, k); \
    break;
    EMEI_AK(EMEI_ACTION_DISCRETE_U8)
    EMEI_AK(EMEI_ACTION_DISCRETE_I32)
    EMEI_AK(EMEI_ACTION_DISCRETE_I64)
    EMEI_AK(EMEI_ACTION_CONTINUOUS_F32)
    EMEI_AK(EMEI_ACTION_CONTINUOUS_F64)
#undef EMEI_AK
  }
}

}  // namespace emei
