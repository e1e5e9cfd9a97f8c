// Fused T-step rollout for the cart-pole / analytic inverted-pendulum family (float32).
//
// Replaces the canonical caller of step(): the collection loop of zoo/util.py:33-93
//     obs = env.reset(); while not done: a = policy(obs) | action_space.sample(); step; record; ...
// for a batch of independent envs, with gym's TimeLimit (register_env.py max_episode_steps: truncated when
// the episode step count reaches the limit) and auto-reset (zoo/util.py:52-54) done in-kernel.
//
// One thread owns one env for the whole horizon: the 4-scalar state, the episode step counter and the
// episode return live in registers across all T steps, so HBM is touched only for what the caller asks
// for -- the action stream (teacher-forced policy) and/or the transition records in the reference's
// dataset layout (observations, next_observations, actions, rewards, dones, timeouts; zoo/util.py:62-67).
// With the built-in uniform random policy and no records a rollout moves 24 bytes per env in total.
//
// Arithmetic per step is the step kernel's (cartpole_f32.cuh): a rollout equals T calls of
// emei_cartpole_step_f32 plus the bookkeeping, bit for bit (tests/test_gpu_parity.py).
#pragma once
#include "cartpole_f32.cuh"

namespace emei {

constexpr uint32_t kPurposeRolloutAction = 4;

struct RolloutConsts {
  int horizon, max_episode_steps, auto_reset, random_policy, init_kind /*0 uniform, 1 gaussian*/, init_pi_column;
  unsigned long long seed_reset, seed_action, env_offset, t0;
  double init_low, init_high, mean[4], sigma[4];
  float act_low, act_high;
};

// same arithmetic as init_uniform_kernel / init_gaussian_kernel (kernels.cuh), one env row
__device__ __noinline__ float4 rollout_init_state(const RolloutConsts& r, unsigned long long env, unsigned long long seed) {
  float v[4];
  if (r.init_kind == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t w[4];
      Philox::generate(seed, env, static_cast<uint32_t>(c >> 1), 1u /*kPurposeUniform*/, w);
      const double u = (c & 1) ? u01_from_bits(w[2], w[3]) : u01_from_bits(w[0], w[1]);
      double x = __dadd_rn(r.init_low, __dmul_rn(r.init_high - r.init_low, u));
      if (c == r.init_pi_column) x = __dadd_rn(x, 3.141592653589793238462643383279502884);
      v[c] = static_cast<float>(x);
    }
  } else {
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      uint32_t w[4];
      Philox::generate(seed, env, static_cast<uint32_t>(pr), 2u /*kPurposeGaussian*/, w);
      const double u1 = 1.0 - u01_from_bits(w[0], w[1]);
      const double u2 = u01_from_bits(w[2], w[3]);
      const double rad = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      v[2 * pr] = static_cast<float>(__dadd_rn(r.mean[2 * pr], __dmul_rn(r.sigma[2 * pr], __dmul_rn(rad, cs))));
      v[2 * pr + 1] = static_cast<float>(__dadd_rn(r.mean[2 * pr + 1], __dmul_rn(r.sigma[2 * pr + 1], __dmul_rn(rad, sn))));
    }
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

__device__ __forceinline__ double block_sum_double(double v, double* smem /*[kBlock/32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double t = lane < kBlock / 32 ? smem[lane] : 0.0;
  return warp_sum(t);  // valid in every lane of every warp
}

__device__ __forceinline__ float ip_wrap(float th) { return wrap_pi_f32(th); }

template <bool IP, int AK, int FR, bool RECORD>
__global__ void __launch_bounds__(kBlock, 4)
    cartpole_rollout_f32_kernel(float4* state_io, int32_t* ep_step_io, float* ep_return_io, int32_t* ep_index_io,
                                const void* __restrict__ actions, float4* __restrict__ rec_obs,
                                float4* __restrict__ rec_next, void* __restrict__ rec_act, float* __restrict__ rec_rew,
                                uint8_t* __restrict__ rec_done, uint8_t* __restrict__ rec_timeout, double* stats,
                                uint32_t n, const CartPoleF32Consts k, const RolloutConsts r) {
  using ActT = typename ActionStorage<AK>::type;
  constexpr bool kDiscrete = AK <= EMEI_ACTION_DISCRETE_I64;
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  const bool live = i < n;
  const bool swingup_ip = IP && (k.variant == EMEI_IP_REBOUND_SWINGUP || k.variant == EMEI_IP_BOUNDARY_SWINGUP);
  const float sgn = swingup_ip ? -1.0f : 1.0f;
  const ActT* act_in = static_cast<const ActT*>(actions);
  ActT* act_out = static_cast<ActT*>(rec_act);
  const unsigned long long env = r.env_offset + i;

  float r_sum = 0.f, fin_ret = 0.f;
  unsigned n_term = 0, n_trunc = 0, n_fin = 0, fin_len = 0;
  pdl_trigger();
  pdl_wait();
  if (live) {
    float4 y = state_io[i];
    int ep_step = ep_step_io[i];
    float ep_ret = ep_return_io[i];
    int ep_idx = ep_index_io[i];
    uint32_t w[4] = {0, 0, 0, 0};
    float a_next = 0.f;
    if (!r.random_policy) a_next = static_cast<float>(__ldg(act_in + i));
    for (int t = 0; t < r.horizon; ++t) {
      // ---- policy
      float a;
      if (r.random_policy) {  // env.action_space.sample() (zoo/util.py:57): Discrete(2) bit / Box uniform
        const unsigned long long tg = r.t0 + static_cast<unsigned long long>(t);
        if (t == 0 || (tg & 3ull) == 0) Philox::generate(r.seed_action, env, static_cast<uint32_t>(tg >> 2), kPurposeRolloutAction, w);
        const uint32_t word = w[tg & 3ull];
        if constexpr (kDiscrete)
          a = static_cast<float>(word & 1u);
        else
          a = fmaf(r.act_high - r.act_low, static_cast<float>(word >> 8) * (1.0f / 16777216.0f), r.act_low);
      } else {
        a = a_next;
        if (t + 1 < r.horizon) a_next = static_cast<float>(__ldg(act_in + static_cast<size_t>(t + 1) * n + i));
      }
      const size_t rec = static_cast<size_t>(t) * n + i;
      if constexpr (RECORD) {
        rec_obs[rec] = IP ? make_float4(y.x, ip_wrap(y.y), y.z, y.w) : y;
        act_out[rec] = static_cast<ActT>(a);
      }
      // ---- dynamics (identical to cartpole_step_f32_kernel)
      float f_mt;
      if constexpr (!IP) {
        float force;
        if constexpr (kDiscrete)
          force = a == 1.0f ? k.force_mag : -k.force_mag;
        else
          force = k.force_mag * a;
        f_mt = force * k.k.inv_mt;
      } else {
        float ctrl = a;
        ctrl = ctrl < k.ctrl_low ? k.ctrl_low : (ctrl > k.ctrl_high ? k.ctrl_high : ctrl);
        f_mt = (k.force_mag * ctrl) * k.k.inv_mt;
      }
      const float4 y0 = y;
      const float th_max = integrate<IP, FR, false>(y, f_mt, sgn, k);
      const bool sane = th_max <= f32::kSinCosSaneMax;
      if (!sane) {
        y = y0;
        integrate<IP, FR, true>(y, f_mt, sgn, k);
      }
      float rew;
      bool notdone;
      float4 obs = y;
      if constexpr (!IP) {
        if (k.variant == EMEI_CARTPOLE_SWINGUP) {
          const float cth = sane ? f32::cos_core(y.z) : cosf(y.z);
          rew = fmaf(cth, 0.5f, 0.5f);
          notdone = fabsf(y.x) < k.x_thr;
        } else {
          rew = 1.0f;
          notdone = (fabsf(y.z) < k.th_thr) && (fabsf(y.x) < k.x_thr);
        }
      } else {
        const float th_obs = ip_wrap(y.y);
        obs.y = th_obs;
        const bool finite = isfinite(y.x) && isfinite(th_obs) && isfinite(y.z) && isfinite(y.w);
        const float cy = f32::cos_core(th_obs);
        const bool in_rail = (k.x_left < y.x) && (y.x < k.x_right);
        switch (k.variant) {
          case EMEI_IP_REBOUND_BALANCING: rew = 1.0f; notdone = (cy >= 0.9f) && finite; break;
          case EMEI_IP_BOUNDARY_BALANCING: rew = 1.0f; notdone = (cy >= 0.0f) && in_rail && finite; break;
          case EMEI_IP_REBOUND_SWINGUP: rew = fmaf(cy, -0.5f, 0.5f); notdone = finite; break;
          default: rew = fmaf(cy, -0.5f, 0.5f); notdone = in_rail && finite; break;
        }
      }
      // ---- TimeLimit + bookkeeping (zoo/util.py:58-73; gym TimeLimit: truncated = elapsed >= max)
      ep_step += 1;
      ep_ret += rew;
      const bool terminated = !notdone;
      const bool truncated = r.max_episode_steps > 0 && ep_step >= r.max_episode_steps;
      const bool done = terminated || truncated;
      if constexpr (RECORD) {
        rec_next[rec] = obs;
        rec_rew[rec] = rew;
        rec_done[rec] = done ? 1 : 0;
        rec_timeout[rec] = truncated ? 1 : 0;
      }
      r_sum += rew;
      n_term += terminated ? 1u : 0u;
      n_trunc += truncated ? 1u : 0u;
      if (done && r.auto_reset) {
        n_fin += 1u;
        fin_ret += ep_ret;
        fin_len += static_cast<unsigned>(ep_step);
        ep_idx += 1;
        y = rollout_init_state(r, env, r.seed_reset + static_cast<unsigned long long>(ep_idx) * 0xD1B54A32D192ED03ull);
        ep_step = 0;
        ep_ret = 0.f;
      }
    }
    state_io[i] = y;
    ep_step_io[i] = ep_step;
    ep_return_io[i] = ep_ret;
    ep_index_io[i] = ep_idx;
  }
  if (stats != nullptr) {  // uniform across the grid
    __shared__ double s_red[kBlock / 32];
    const double vals[6] = {static_cast<double>(r_sum), static_cast<double>(n_term), static_cast<double>(n_trunc),
                            static_cast<double>(n_fin), static_cast<double>(fin_ret), static_cast<double>(fin_len)};
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double tot = block_sum_double(vals[j], s_red);
      if (threadIdx.x == 0) atomicAdd(&stats[j], tot);
    }
  }
}

struct RolloutBuffers {
  float* state_io; int32_t* ep_step_io; float* ep_return_io; int32_t* ep_index_io; const void* actions;
  float* rec_obs; float* rec_next; void* rec_act; float* rec_rew; uint8_t* rec_done; uint8_t* rec_timeout; double* stats;
};

template <bool IP, int AK, int FR>
inline void launch_rollout(const RolloutBuffers& b, int64_t n, const CartPoleF32Consts& k, const RolloutConsts& r, cudaStream_t s) {
  const int grid = grid_for(n, kBlock);
  float4* st = reinterpret_cast<float4*>(b.state_io);
  float4* ro = reinterpret_cast<float4*>(b.rec_obs);
  float4* rn = reinterpret_cast<float4*>(b.rec_next);
  if (b.rec_obs != nullptr)
    launch_pdl(cartpole_rollout_f32_kernel<IP, AK, FR, true>, grid, kBlock, s, st, b.ep_step_io, b.ep_return_io, b.ep_index_io,
               b.actions, ro, rn, b.rec_act, b.rec_rew, b.rec_done, b.rec_timeout, b.stats, static_cast<uint32_t>(n), k, r);
  else
    launch_pdl(cartpole_rollout_f32_kernel<IP, AK, FR, false>, grid, kBlock, s, st, b.ep_step_io, b.ep_return_io, b.ep_index_io,
               b.actions, ro, rn, b.rec_act, b.rec_rew, b.rec_done, b.rec_timeout, b.stats, static_cast<uint32_t>(n), k, r);
}

template <bool IP, int FR>
inline void launch_rollout_ak(int ak, const RolloutBuffers& b, int64_t n, const CartPoleF32Consts& k, const RolloutConsts& r, cudaStream_t s) {
  switch (ak) {
    case EMEI_ACTION_DISCRETE_U8: launch_rollout<IP, EMEI_ACTION_DISCRETE_U8, FR>(b, n, k, r, s); break;
    case EMEI_ACTION_DISCRETE_I32: launch_rollout<IP, EMEI_ACTION_DISCRETE_I32, FR>(b, n, k, r, s); break;
    case EMEI_ACTION_DISCRETE_I64: launch_rollout<IP, EMEI_ACTION_DISCRETE_I64, FR>(b, n, k, r, s); break;
    case EMEI_ACTION_CONTINUOUS_F32: launch_rollout<IP, EMEI_ACTION_CONTINUOUS_F32, FR>(b, n, k, r, s); break;
    default: launch_rollout<IP, EMEI_ACTION_CONTINUOUS_F64, FR>(b, n, k, r, s); break;
  }
}

}  // namespace emei
