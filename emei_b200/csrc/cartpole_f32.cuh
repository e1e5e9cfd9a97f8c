// float32 (tolerance-mode) cart-pole / analytic inverted-pendulum step: the headline kernel.
//
// Replaces BaseControlEnv.step (base_control.py:61-83), ODE_approximation's forward-Euler loop
// (base_control.py:160-164), BaseCartPoleEnv._dsdt (cartpole.py:48-60) and the reward/terminal of
// cartpole.py:124-129,145-151 / inverted_pendulum.py:73-79,103-111,139-146,174-183 for a batch.
//
// Shape: persistent grid (148 SMs x 4-6 CTAs of 256 threads), grid-stride over envs, one env per
// thread per iteration.  Each thread keeps the inputs of its next S (2-4) envs in flight with
// cp.async (LDGSTS) into a private column of a shared-memory ring -- no registers held, no barriers --
// so HBM streams continuously under the ~40 instructions/sub-step of math.  freq_rate 1 and
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
// Prefetch loads are written as volatile PTX so the compiler keeps them at the TOP of the loop body:
// left to itself nvcc sinks the next env's loads below the unrolled sub-steps (to shorten live
// ranges), which exposes the full HBM latency once per env (profiles/r01: 11.6 vs 11.3 us).
__device__ __forceinline__ float4 ld_state_early(const float4* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

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

// ---- asynchronous staging (LDGSTS): global -> shared without passing through registers ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

template <int AK>
struct ActionStorage {  // element type of the action array
  using type = float;
};
template <> struct ActionStorage<EMEI_ACTION_DISCRETE_U8> { using type = uint8_t; };
template <> struct ActionStorage<EMEI_ACTION_DISCRETE_I32> { using type = int32_t; };
template <> struct ActionStorage<EMEI_ACTION_DISCRETE_I64> { using type = long long; };
template <> struct ActionStorage<EMEI_ACTION_CONTINUOUS_F64> { using type = double; };


template <bool IP, int AK, int FR, int MINB, bool HAS_OBS, int S, int DBG = 0>
__global__ void __launch_bounds__(kBlock, MINB)
    cartpole_step_f32_kernel(const float4* state_in, float4* state_out, float4* obs_out,
                             const void* __restrict__ action, float* __restrict__ reward, uint8_t* __restrict__ done,
                             double* stats, uint32_t n, const CartPoleF32Consts k) {
  // 32-bit env index: the launcher splits batches above 2^31 - 2^20 envs (never in practice: that is
  // 32 GiB of float32 state)
  using ActT = typename ActionStorage<AK>::type;
  constexpr bool kStageAction = sizeof(ActT) >= 4;  // cp.async moves 4/8/16 bytes; uint8 actions ride in a register
  __shared__ float4 s_state[S][kBlock];
  __shared__ ActT s_act[kStageAction ? S : 1][kStageAction ? kBlock : 1];
  const uint32_t tid = threadIdx.x;
  const uint32_t stride = gridDim.x * kBlock;
  uint32_t i = blockIdx.x * kBlock + tid;
  float r_acc = 0.0f;
  unsigned d_cnt = 0;
  constexpr bool kDiscrete = AK <= EMEI_ACTION_DISCRETE_I64;
  const bool swingup_ip = IP && (k.variant == EMEI_IP_REBOUND_SWINGUP || k.variant == EMEI_IP_BOUNDARY_SWINGUP);
  const float sgn = swingup_ip ? -1.0f : 1.0f;
  const ActT* act = static_cast<const ActT*>(action);

  // one env's inputs -> ring slot `slot` (each thread only ever touches its own column of the ring,
  // so completion of its own cp.async groups is all the synchronisation the ring needs)
  auto stage_in = [&](int slot, uint32_t idx) {
    if (idx < n && DBG != 2) {
      cp_async<16>(&s_state[slot][tid], state_in + idx);
      if constexpr (kStageAction) cp_async<sizeof(ActT)>(&s_act[slot][tid], act + idx);
    }
    cp_async_commit();
  };

  pdl_trigger();  // let the next step kernel of the rollout be staged behind this one
  pdl_wait();     // the previous kernel in the stream (the step that wrote state_in) has completed
#pragma unroll
  for (int d = 0; d < S; ++d) stage_in(d, i + d * stride);  // (i + d*stride cannot wrap: n <= 2^31 - 2^20)
  [[maybe_unused]] float a_reg = 0.f;
  if constexpr (!kStageAction)
    if (i < n) a_reg = static_cast<float>(__ldg(act + i));

  int slot = 0;
  while (i < n) {
    cp_async_wait<S - 1>();  // the oldest group (this env) has landed
    float4 y = s_state[slot][tid];
    // DBG != 0 exists only for tools/kbench (1: stores elided, 2: loads elided); the library uses 0
    if constexpr (DBG == 2) y = make_float4(1e-6f * i, 0.5f, 2e-6f * i, -1.0f);  // kbench: no dependence on loads
    float a;
    if constexpr (kStageAction) {
      a = static_cast<float>(s_act[slot][tid]);
    } else {
      a = a_reg;
      const uint32_t i_next = i + stride;
      if (i_next < n) a_reg = static_cast<float>(__ldg(act + i_next));
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
    stage_in(slot, i + S * stride);  // refill this slot (its previous content is consumed: y depends on it)
    slot = slot + 1 == S ? 0 : slot + 1;
    const bool sane = th_max <= f32::kSinCosSaneMax;
    if (!sane) {
      y = state_in[i];
      integrate<IP, FR, true>(y, f_mt, sgn, k);
    }
    if (DBG != 1 || y.x == 1234.5f) state_out[i] = y;
    float rew;
    bool notdone;
    if constexpr (!IP) {
      if constexpr (HAS_OBS) obs_out[i] = y;
      if (k.variant == EMEI_CARTPOLE_SWINGUP) {
        const float cth = sane ? f32::cos_core(y.z) : cosf(y.z);
        rew = fmaf(cth, 0.5f, 0.5f);     // cartpole.py:149-151
        notdone = fabsf(y.x) < k.x_thr;  // cartpole.py:145-147
      } else {
        rew = 1.0f;                                                      // cartpole.py:128-129
        notdone = (fabsf(y.z) < k.th_thr) && (fabsf(y.x) < k.x_thr);  // cartpole.py:124-126
      }
    } else {
      // observation: theta wrapped to [-pi, pi) (inverted_pendulum.py:45-49)
      const float th_obs = wrap_pi_f32(y.y);
      if constexpr (HAS_OBS) obs_out[i] = make_float4(y.x, th_obs, y.z, y.w);
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
    const unsigned d = notdone ? 0u : 1u;
    if (DBG != 1 || rew == 1234.5f) {
      reward[i] = rew;
      done[i] = static_cast<uint8_t>(d);
    }
    r_acc += rew;
    d_cnt += d;
    i += stride;
  }
  cp_async_wait<0>();
  block_stats_accumulate_counts(stats, static_cast<double>(r_acc), d_cnt);
}

// Launch shape (tools/kbench, B200, 2^20 envs, freq_rate 4: 10.9 us/launch vs 11.1-13.8 for the other
// combinations): unrolled sub-steps want registers (4 CTAs/SM -> 64 regs, 4-deep ring); the run-time
// loop is happiest at 6 CTAs/SM with a 2-deep ring.
template <int FR>
struct CartPoleShape {
  static constexpr int kMinBlocks = FR > 0 ? 4 : 6;
  static constexpr int kStages = FR > 0 ? 4 : 2;
};
constexpr int64_t kCartPoleMaxLaunch = (1ll << 31) - (1ll << 20);  // envs per launch (32-bit index)

template <bool IP, int FR>
inline void launch_cartpole_f32(int ak, cudaStream_t s, const float* state_in, float* state_out, float* obs_out,
                                const void* action, int action_bytes, float* reward, uint8_t* done, double* stats,
                                int64_t n, const CartPoleF32Consts& k) {
  constexpr int MB = CartPoleShape<FR>::kMinBlocks, ST = CartPoleShape<FR>::kStages;
  for (int64_t off = 0; off < n; off += kCartPoleMaxLaunch) {
    const int64_t m = n - off < kCartPoleMaxLaunch ? n - off : kCartPoleMaxLaunch;
    const int grid = persistent_grid(m, kBlock, MB);
    const float4* in4 = reinterpret_cast<const float4*>(state_in) + off;
    float4* out4 = reinterpret_cast<float4*>(state_out) + off;
    float4* obs4 = obs_out ? reinterpret_cast<float4*>(obs_out) + off : nullptr;
    const void* act = static_cast<const char*>(action) + off * action_bytes;
    switch (ak) {
#define EMEI_AK(A)                                                                                                  \
  case A:                                                                                                           \
    if (obs4 != nullptr)                                                                                            \
      launch_pdl(cartpole_step_f32_kernel<IP, A, FR, MB, true, ST>, grid, kBlock, s, in4, out4, obs4, act, reward + off, \
                 done + off, stats, static_cast<uint32_t>(m), k);                                                   \
    else                                                                                                            \
      launch_pdl(cartpole_step_f32_kernel<IP, A, FR, MB, false, ST>, grid, kBlock, s, in4, out4, obs4, act, reward + off, \
                 done + off, stats, static_cast<uint32_t>(m), k);                                                   \
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
  const bool ip = p.variant > EMEI_CARTPOLE_SWINGUP;
  const int ab = action_kind_bytes(p.action_kind);
#define EMEI_GO(IPV, FRV) \
  launch_cartpole_f32<IPV, FRV>(p.action_kind, s, state_in, state_out, obs_out, action, ab, reward, done, stats, n, k)
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
