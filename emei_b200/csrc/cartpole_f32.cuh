// float32 (tolerance-mode) cart-pole / analytic inverted-pendulum step: the headline kernel.
//
// Replaces BaseControlEnv.step (base_control.py:61-83), ODE_approximation's forward-Euler loop
// (base_control.py:160-164), BaseCartPoleEnv._dsdt (cartpole.py:48-60) and the reward/terminal of
// cartpole.py:124-129,145-151 / inverted_pendulum.py:73-79,103-111,139-146,174-183 for a batch.
//
// Shape: persistent grid (<= 148 SMs x 8 CTAs of 256 threads), grid-stride over envs, one env per
// thread per iteration, the NEXT env's 20 bytes prefetched into registers before the current one
// is integrated, so loads stay in flight under the ~40 instructions/sub-step of math.  freq_rate 1 and
// 4 are compiled unrolled (constants hoisted into registers); other values use a run-time loop.  The 4-scalar
// state lives in registers across all freq_rate sub-steps: HBM sees one 128-bit load, one 128-bit
// store (+ action 4 B, reward 4 B, done 1 B) per env step = 41 bytes.  Per-thread reward / done
// partials are reduced once per thread (warp shuffles -> one atomic pair per CTA).
#pragma once
#include "common.cuh"
#include "f32math.cuh"

namespace emei {

struct CartPoleF32Consts {
  f32::CartPoleK k;
  float force_mag, x_thr, th_thr, x_left, x_right, ctrl_low, ctrl_high;
  int freq_rate, variant;
};

inline CartPoleF32Consts make_cartpole_f32_consts(const emei_cartpole_params& p) {
  CartPoleF32Consts c;
  c.k.g = static_cast<float>(p.gravity);
  c.k.kpm = static_cast<float>(p.pole_mass_length / p.total_mass);
  c.k.inv_mt = static_cast<float>(1.0 / p.total_mass);
  c.k.den0 = static_cast<float>(p.length * (4.0 / 3.0));
  c.k.den1 = static_cast<float>(p.length * p.mass_pole / p.total_mass);
  c.k.dt = static_cast<float>(p.dt);
  c.force_mag = static_cast<float>(p.force_mag);
  c.x_thr = static_cast<float>(p.x_threshold);
  c.th_thr = static_cast<float>(p.theta_threshold);
  c.x_left = static_cast<float>(p.x_left);
  c.x_right = static_cast<float>(p.x_right);
  c.ctrl_low = static_cast<float>(p.ctrl_low);
  c.ctrl_high = static_cast<float>(p.ctrl_high);
  c.freq_rate = p.freq_rate;
  c.variant = p.variant;
  return c;
}

// raw action value as float (discrete kinds: the integer itself; only an exact 1 means "push right")
template <int AK>
__device__ __forceinline__ float load_action_f32(const void* __restrict__ action, int64_t i) {
  if constexpr (AK == EMEI_ACTION_DISCRETE_U8)
    return static_cast<float>(__ldg(static_cast<const uint8_t*>(action) + i));
  else if constexpr (AK == EMEI_ACTION_DISCRETE_I32)
    return static_cast<float>(__ldg(static_cast<const int32_t*>(action) + i));
  else if constexpr (AK == EMEI_ACTION_DISCRETE_I64)
    return static_cast<float>(__ldg(static_cast<const long long*>(action) + i));
  else if constexpr (AK == EMEI_ACTION_CONTINUOUS_F32)
    return __ldg(static_cast<const float*>(action) + i);
  else
    return static_cast<float>(__ldg(static_cast<const double*>(action) + i));
}

// One env step for the state in registers.  FR > 0: compile-time sub-step count (unrolled);
// FR == 0: run-time k.freq_rate.  LIBM selects the guarded slow sincos (cold path).
template <bool IP, int FR, bool LIBM>
__device__ __forceinline__ float integrate(float4& y, float f_mt, float sgn, const CartPoleF32Consts& k) {
  const int fr = FR > 0 ? FR : k.freq_rate;
  float th_max = fabsf(IP ? y.y : y.z);
#pragma unroll
  for (int sub = 0; sub < fr; ++sub) {
    if constexpr (!IP) {
      f32::cartpole_substep<LIBM>(y.x, y.y, y.z, y.w, f_mt, 1.0f, k.k);  // [x, x_dot, theta, theta_dot]
      th_max = fmaxf(th_max, fabsf(y.z));
    } else {
      f32::cartpole_substep<LIBM>(y.x, y.z, y.y, y.w, f_mt, sgn, k.k);  // [x, theta, v, omega]
      th_max = fmaxf(th_max, fabsf(y.y));
    }
  }
  return th_max;
}

template <bool IP, int AK, int FR, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
    cartpole_step_f32_kernel(const float4* state_in, float4* state_out, float4* obs_out,
                             const void* __restrict__ action, float* __restrict__ reward, uint8_t* __restrict__ done,
                             double* stats, uint32_t n, const CartPoleF32Consts k) {
  // 32-bit env index: the launcher splits batches above 2^31 - 2^20 envs (never in practice: that is
  // 32 GiB of float32 state)
  const uint32_t stride = gridDim.x * kBlock;
  uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  float r_acc = 0.0f;
  unsigned d_cnt = 0;
  constexpr bool kDiscrete = AK <= EMEI_ACTION_DISCRETE_I64;
  const bool swingup_ip = IP && (k.variant == EMEI_IP_REBOUND_SWINGUP || k.variant == EMEI_IP_BOUNDARY_SWINGUP);
  const float sgn = swingup_ip ? -1.0f : 1.0f;

  float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
  float a = 0.f;
  if (i < n) {
    y = state_in[i];
    a = load_action_f32<AK>(action, i);
  }
  while (i < n) {
    const uint32_t i_next = i + stride;
    float4 y_next = make_float4(0.f, 0.f, 0.f, 0.f);
    float a_next = 0.f;
    if (i_next < n) {  // prefetch: in flight while this env integrates
      y_next = state_in[i_next];
      a_next = load_action_f32<AK>(action, i_next);
    }
    float f_mt;
    if constexpr (!IP) {
      float force;
      if constexpr (kDiscrete)
        force = a == 1.0f ? k.force_mag : -k.force_mag;  // cartpole.py:121-122,142-143
      else
        force = k.force_mag * a;  // continuous: force_mag * action[0]; held over the sub-steps (cartpole.py:60)
      f_mt = force * k.k.inv_mt;
    } else {
      float ctrl = a;
      ctrl = ctrl < k.ctrl_low ? k.ctrl_low : (ctrl > k.ctrl_high ? k.ctrl_high : ctrl);  // mj_step clamps ctrl
      f_mt = (k.force_mag * ctrl) * k.k.inv_mt;                                          // gear * ctrl
    }
    // the unguarded sincos is valid while |theta| stays below kSinCosSaneMax; otherwise (or NaN) redo
    // this env from its stored state with the libm path.  Cold: float32 theta is meaningless there.
    const float th_max = integrate<IP, FR, false>(y, f_mt, sgn, k);
    if (!(th_max <= f32::kSinCosSaneMax)) {
      y = state_in[i];
      integrate<IP, FR, true>(y, f_mt, sgn, k);
    }
    state_out[i] = y;
    float rew;
    bool notdone;
    if constexpr (!IP) {
      if (obs_out != nullptr) obs_out[i] = y;
      if (k.variant == EMEI_CARTPOLE_SWINGUP) {
        rew = fmaf(f32::cos_fast(y.z), 0.5f, 0.5f);  // cartpole.py:149-151
        notdone = fabsf(y.x) < k.x_thr;               // cartpole.py:145-147
      } else {
        rew = 1.0f;                                                      // cartpole.py:128-129
        notdone = (fabsf(y.z) < k.th_thr) && (fabsf(y.x) < k.x_thr);  // cartpole.py:124-126
      }
    } else {
      // observation: theta wrapped to [-pi, pi) (inverted_pendulum.py:45-49)
      const float th_obs = py_mod(y.y + 3.14159265358979323846f, 6.28318530717958647692f) - 3.14159265358979323846f;
      if (obs_out != nullptr) obs_out[i] = make_float4(y.x, th_obs, y.z, y.w);
      const bool finite = isfinite(y.x) && isfinite(th_obs) && isfinite(y.z) && isfinite(y.w);
      const float cy = f32::cos_core(th_obs);  // |th_obs| <= pi (NaN stays NaN)
      const bool in_rail = (k.x_left < y.x) && (y.x < k.x_right);
      switch (k.variant) {
        case EMEI_IP_REBOUND_BALANCING:  // inverted_pendulum.py:73-79
          rew = 1.0f;
          notdone = (cy >= 0.9f) && finite;
          break;
        case EMEI_IP_BOUNDARY_BALANCING:  // :103-111
          rew = 1.0f;
          notdone = (cy >= 0.0f) && in_rail && finite;
          break;
        case EMEI_IP_REBOUND_SWINGUP:  // :139-146
          rew = fmaf(cy, -0.5f, 0.5f);
          notdone = finite;
          break;
        default:  // EMEI_IP_BOUNDARY_SWINGUP :174-183
          rew = fmaf(cy, -0.5f, 0.5f);
          notdone = in_rail && finite;
          break;
      }
    }
    reward[i] = rew;
    done[i] = notdone ? 0 : 1;
    r_acc += rew;
    d_cnt += notdone ? 0u : 1u;
    y = y_next;
    a = a_next;
    i = i_next;
  }
  block_stats_accumulate_counts(stats, static_cast<double>(r_acc), d_cnt);
}

constexpr int kCartPoleMinBlocks = 6;                       // resident CTAs per SM the kernel is compiled for
constexpr int64_t kCartPoleMaxLaunch = (1ll << 31) - (1ll << 20);  // envs per launch (32-bit index)

template <bool IP, int FR>
inline void launch_cartpole_f32(int ak, int grid, cudaStream_t s, const float* state_in, float* state_out,
                                float* obs_out, const void* action, int action_bytes, float* reward, uint8_t* done,
                                double* stats, int64_t n, const CartPoleF32Consts& k) {
  for (int64_t off = 0; off < n; off += kCartPoleMaxLaunch) {
    const int64_t m = n - off < kCartPoleMaxLaunch ? n - off : kCartPoleMaxLaunch;
    const float4* in4 = reinterpret_cast<const float4*>(state_in) + off;
    float4* out4 = reinterpret_cast<float4*>(state_out) + off;
    float4* obs4 = obs_out ? reinterpret_cast<float4*>(obs_out) + off : nullptr;
    const void* act = static_cast<const char*>(action) + off * action_bytes;
    switch (ak) {
#define EMEI_AK(A)                                                                                     \
  case A:                                                                                              \
    cartpole_step_f32_kernel<IP, A, FR, kCartPoleMinBlocks>                                            \
        <<<grid, kBlock, 0, s>>>(in4, out4, obs4, act, reward + off, done + off, stats, static_cast<uint32_t>(m), k); \
    break;
      EMEI_AK(EMEI_ACTION_DISCRETE_U8)
      EMEI_AK(EMEI_ACTION_DISCRETE_I32)
      EMEI_AK(EMEI_ACTION_DISCRETE_I64)
      EMEI_AK(EMEI_ACTION_CONTINUOUS_F32)
      EMEI_AK(EMEI_ACTION_CONTINUOUS_F64)
#undef EMEI_AK
    }
  }
}

inline int action_kind_bytes(int ak) {
  switch (ak) {
    case EMEI_ACTION_DISCRETE_U8: return 1;
    case EMEI_ACTION_DISCRETE_I32: return 4;
    case EMEI_ACTION_DISCRETE_I64: return 8;
    case EMEI_ACTION_CONTINUOUS_F32: return 4;
    default: return 8;
  }
}

inline void cartpole_step_f32_dispatch(const float* state_in, float* state_out, float* obs_out, const void* action,
                                       float* reward, uint8_t* done, double* stats, int64_t n,
                                       const emei_cartpole_params& p, cudaStream_t s) {
  const CartPoleF32Consts k = make_cartpole_f32_consts(p);
  const int64_t per_launch = n < kCartPoleMaxLaunch ? n : kCartPoleMaxLaunch;
  const int grid = persistent_grid(per_launch, kBlock, kCartPoleMinBlocks);
  const bool ip = p.variant > EMEI_CARTPOLE_SWINGUP;
  const int ab = action_kind_bytes(p.action_kind);
#define EMEI_GO(IPV, FRV) \
  launch_cartpole_f32<IPV, FRV>(p.action_kind, grid, s, state_in, state_out, obs_out, action, ab, reward, done, stats, n, k)
  if (!ip) {
    if (p.freq_rate == 1) EMEI_GO(false, 1);
    else if (p.freq_rate == 4) EMEI_GO(false, 4);
    else EMEI_GO(false, 0);
  } else {
    if (p.freq_rate == 1) EMEI_GO(true, 1);
    else if (p.freq_rate == 4) EMEI_GO(true, 4);
    else EMEI_GO(true, 0);
  }
#undef EMEI_GO
}

}  // namespace emei
