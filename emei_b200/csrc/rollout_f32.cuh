// Fused T-step rollouts (float32): cart-pole / analytic inverted-pendulum family and charged ball.
//
// Replaces the canonical caller of step(): the collection loop of zoo/util.py:33-93
//     obs = env.reset(); while not done: a = policy(obs) | action_space.sample(); step; record; ...
// for a batch of independent envs, with gym's TimeLimit (register_env.py max_episode_steps: truncated when
// the episode step count reaches the limit) and auto-reset (zoo/util.py:52-54) done in-kernel.
//
// One thread owns one env for the whole horizon: the env's state, the episode step counter and the
// episode return live in registers across all T steps, so HBM is touched only for what the caller asks
// for -- the action stream (teacher-forced policy) and/or the transition records in the reference's
// dataset layout (observations, next_observations, actions, rewards, dones, timeouts; zoo/util.py:62-67).
// With the built-in uniform random policy and no records a rollout moves < 100 bytes per env in total.
//
// Arithmetic per step is the step kernels' (cartpole_f32.cuh / charged_ball_f32.cuh): a rollout equals T
// calls of emei_*_step_f32 plus the bookkeeping, bit for bit (tests/test_gpu_parity.py).
//
// The kernel is written once over a `Dyn` policy (the env family): Dyn::Buffers (state arrays), Dyn::Regs
// (one env in registers), load/store/observation/step/reset.
#pragma once
#include "cartpole_f32.cuh"
#include "charged_ball_f32.cuh"

namespace emei {

constexpr uint32_t kPurposeRolloutAction = 4;

struct RolloutConsts {
  int horizon, max_episode_steps, auto_reset, random_policy, init_kind /*0 uniform, 1 gaussian, 2 charged ball*/, init_pi_column;
  unsigned long long seed_reset, seed_action, env_offset, t0;
  double init_low, init_high, mean[4], sigma[4];
  float act_low, act_high;
};

constexpr uint32_t kPurposeRolloutReset = 5;

// In-rollout reset sample of one env row.
// Uniform family (cartpole.py:131-132,153-156: U(low, high) per coordinate, + pi on the swing-up angle): the lean
// float32 sampler of the rollout kernels -- ONE Philox block per reset, coordinate c = fma(high - low, u_c, low)
// with u_c the top 24 bits of word c.  A reset is a divergent branch that a third of all warp-steps take under
// the random policy, so its length is paid by the whole warp: the 53-bit double construction of the init kernels
// (two blocks, four double conversions) cost ~360 warp instructions per event, this one ~110
// (mirror: oracle/rollout_oracle.py reset_sample_uniform).
// Gaussian family (mujoco_env.py:137-140): the arithmetic of init_gaussian_kernel (kernels.cuh).
static __device__ __noinline__ float4 rollout_init_state(const RolloutConsts& r, unsigned long long env, unsigned long long seed) {
  float v[4];
  if (r.init_kind == 0) {
    uint32_t w[4];
    Philox::generate(seed, env, 0u, kPurposeRolloutReset, w);
    const float lo = static_cast<float>(r.init_low), span = static_cast<float>(r.init_high) - lo;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = fmaf(span, static_cast<float>(w[c] >> 8) * (1.0f / 16777216.0f), lo);
      if (c == r.init_pi_column) v[c] = static_cast<float>(static_cast<double>(v[c]) + 3.141592653589793238462643383279502884);
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

// same arithmetic as init_charged_ball_kernel<float> (kernels.cuh; charged_ball.py:84-94), one env
static __device__ __noinline__ CBRegs rollout_init_charged_ball(float radius, unsigned long long env, unsigned long long seed) {
  uint32_t w[4];
  Philox::generate(seed, env, 0u, 3u /*kPurposeChargedBall*/, w);
  const double theta = __dadd_rn(__dadd_rn(-0.5, u01_from_bits(w[0], w[1])), 3.141592653589793238462643383279502884);
  const double omega = __dadd_rn(-0.5, u01_from_bits(w[2], w[3]));
  CBRegs e;
  e.on = true;
  e.theta = static_cast<float>(theta);
  e.omega = static_cast<float>(omega);
  float s, c;
  sincosf(e.theta, &s, &c);  // libm like the init kernel: the sample's bits do not depend on who draws it
  const float x = s * radius, y = c * radius;
  e.f = make_float4(x, y, e.omega * y, -e.omega * x);
  cb_prepare(e);
  return e;
}

// w[j] for a run-time j without putting the array in local memory
__device__ __forceinline__ uint32_t select_word(const uint32_t (&w)[4], unsigned j) {
  const uint32_t lo = (j & 1u) ? w[1] : w[0], hi = (j & 1u) ? w[3] : w[2];
  return (j & 2u) ? hi : lo;
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

// ------------------------------------------------------------------------------------------------
// env families
// ------------------------------------------------------------------------------------------------
template <bool IP, int AK, int FR>
struct CartPoleDyn {
  using Consts = CartPoleF32Consts;
  using ActT = typename ActionStorage<AK>::type;
  static constexpr bool kDiscrete = AK <= EMEI_ACTION_DISCRETE_I64;
  static constexpr int kMinBlocks = 4;
  struct Buffers {
    float4* state;
  };
  struct Regs {
    float4 y;
  };
  __device__ __forceinline__ static void load(Regs& e, const Buffers& b, uint32_t i) { e.y = b.state[i]; }
  __device__ __forceinline__ static void store(const Regs& e, const Buffers& b, uint32_t i) { b.state[i] = e.y; }
  __device__ __forceinline__ static float4 observation(const Regs& e) {
    return IP ? make_float4(e.y.x, wrap_pi_f32(e.y.y), e.y.z, e.y.w) : e.y;
  }
  // the scalar (one env per thread) form of the arithmetic of cartpole_step_f32_tma_kernel: same bits
  __device__ __forceinline__ static void step(Regs& e, float a, const Consts& k, float& rew, bool& terminated, float4& next_obs) {
    const uint32_t flip = ip_flip(IP, k.variant);
    const float f_mt = action_to_f_mt<IP, AK>(a, k);
    float4 y = e.y;
    bool notdone;
    cartpole_step_one<IP, FR>(y, f_mt, flip, k, rew, notdone, next_obs);
    e.y = y;
    terminated = !notdone;
  }
  __device__ __forceinline__ static void reset(Regs& e, const RolloutConsts& r, const Consts&, unsigned long long env, unsigned long long seed) {
    e.y = rollout_init_state(r, env, seed);
  }
};

template <int AK>
struct ChargedBallDyn {
  using Consts = ChargedBallF32Consts;
  using ActT = typename ActionStorage<AK>::type;
  static constexpr bool kDiscrete = AK <= EMEI_ACTION_DISCRETE_I64;
  static constexpr int kMinBlocks = 4;
  struct Buffers {
    uint8_t* on_circle;
    float2* circle;
    float4* free_state;
  };
  using Regs = CBRegs;
  __device__ __forceinline__ static void load(Regs& e, const Buffers& b, uint32_t i) {
    e.on = b.on_circle[i] != 0;
    const float2 c2 = b.circle[i];
    e.theta = c2.x;
    e.omega = c2.y;
    e.f = b.free_state[i];
    cb_prepare(e);
  }
  __device__ __forceinline__ static void store(const Regs& e, const Buffers& b, uint32_t i) {
    b.on_circle[i] = e.on ? 1 : 0;
    b.circle[i] = make_float2(e.theta, e.omega);
    b.free_state[i] = e.f;
  }
  __device__ __forceinline__ static float4 observation(const Regs& e) { return e.f; }  // charged_ball.py:96-97
  __device__ __forceinline__ static void step(Regs& e, float a, const Consts& k, float& rew, bool& terminated, float4& next_obs) {
    rew = cb_env_step(e, cb_field<AK>(a, k), k);
    terminated = false;  // charged_ball.py:110-111
    next_obs = e.f;
  }
  __device__ __forceinline__ static void reset(Regs& e, const RolloutConsts&, const Consts& k, unsigned long long env, unsigned long long seed) {
    e = rollout_init_charged_ball(k.r, env, seed);
  }
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct RolloutIO {  // everything that is not the env family's own state
  int32_t* ep_step;
  float* ep_return;
  int32_t* ep_index;
  const void* actions;
  float4* rec_obs;
  float4* rec_next;
  void* rec_act;
  float* rec_rew;
  uint8_t* rec_done;
  uint8_t* rec_timeout;
  double* stats;
};

template <class Dyn, bool RECORD>
__global__ void __launch_bounds__(kBlock, Dyn::kMinBlocks)
    rollout_f32_kernel(const typename Dyn::Buffers b, const RolloutIO io, uint32_t n, const typename Dyn::Consts k,
                       const RolloutConsts r) {
  using ActT = typename Dyn::ActT;
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  const bool live = i < n;
  const ActT* __restrict__ act_in = static_cast<const ActT*>(io.actions);
  ActT* __restrict__ act_out = static_cast<ActT*>(io.rec_act);
  const unsigned long long env = r.env_offset + i;

  float r_sum = 0.f, fin_ret = 0.f;
  unsigned n_term = 0, n_trunc = 0, n_fin = 0, fin_len = 0;
  pdl_trigger();
  pdl_wait();
  if (live) {
    typename Dyn::Regs e;
    Dyn::load(e, b, i);
    int ep_step = io.ep_step[i];
    float ep_ret = io.ep_return[i];
    int ep_idx = io.ep_index[i];
    uint32_t w[4] = {0, 0, 0, 0};
    float a_next = 0.f;
    if (!r.random_policy) a_next = static_cast<float>(__ldg(act_in + i));
    for (int t = 0; t < r.horizon; ++t) {
      // ---- policy
      float a;
      if (r.random_policy) {  // env.action_space.sample() (zoo/util.py:57): Discrete(2) bit / Box uniform
        // counter-based stream per (seed_action, env): Discrete(2) consumes ONE bit per step (a 128-bit Philox
        // block lasts 128 steps), Box one 32-bit word per step (top 24 bits -> [low, high))
        const unsigned long long tg = r.t0 + static_cast<unsigned long long>(t);
        if constexpr (Dyn::kDiscrete) {
          if (t == 0 || (tg & 127ull) == 0) Philox::generate(r.seed_action, env, static_cast<uint32_t>(tg >> 7), kPurposeRolloutAction, w);
          const uint32_t word = select_word(w, static_cast<unsigned>(tg >> 5) & 3u);
          a = static_cast<float>((word >> (static_cast<unsigned>(tg) & 31u)) & 1u);
        } else {
          if (t == 0 || (tg & 3ull) == 0) Philox::generate(r.seed_action, env, static_cast<uint32_t>(tg >> 2), kPurposeRolloutAction, w);
          const uint32_t word = select_word(w, static_cast<unsigned>(tg) & 3u);
          a = fmaf(r.act_high - r.act_low, static_cast<float>(word >> 8) * (1.0f / 16777216.0f), r.act_low);
        }
      } else {
        a = a_next;
        if (t + 1 < r.horizon) a_next = static_cast<float>(__ldg(act_in + static_cast<size_t>(t + 1) * n + i));
      }
      const size_t rec = static_cast<size_t>(t) * n + i;
      if constexpr (RECORD) {
        io.rec_obs[rec] = Dyn::observation(e);
        act_out[rec] = static_cast<ActT>(a);
      }
      // ---- dynamics + reward + terminal (the step kernel's arithmetic)
      float rew;
      bool terminated;
      float4 next_obs;
      Dyn::step(e, a, k, rew, terminated, next_obs);
      // ---- TimeLimit + bookkeeping (zoo/util.py:58-73; gym TimeLimit: truncated = elapsed >= max)
      ep_step += 1;
      ep_ret += rew;
      const bool truncated = r.max_episode_steps > 0 && ep_step >= r.max_episode_steps;
      const bool done = terminated || truncated;
      if constexpr (RECORD) {
        io.rec_next[rec] = next_obs;
        io.rec_rew[rec] = rew;
        io.rec_done[rec] = done ? 1 : 0;
        io.rec_timeout[rec] = truncated ? 1 : 0;
      }
      r_sum += rew;
      n_term += terminated ? 1u : 0u;
      n_trunc += truncated ? 1u : 0u;
      if (done && r.auto_reset) {
        n_fin += 1u;
        fin_ret += ep_ret;
        fin_len += static_cast<unsigned>(ep_step);
        ep_idx += 1;
        Dyn::reset(e, r, k, env, r.seed_reset + static_cast<unsigned long long>(ep_idx) * 0xD1B54A32D192ED03ull);
        ep_step = 0;
        ep_ret = 0.f;
      }
    }
    Dyn::store(e, b, i);
    io.ep_step[i] = ep_step;
    io.ep_return[i] = ep_ret;
    io.ep_index[i] = ep_idx;
  }
  if (io.stats != nullptr) {  // uniform across the grid
    __shared__ double s_red[kBlock / 32];
    const double vals[6] = {static_cast<double>(r_sum), static_cast<double>(n_term), static_cast<double>(n_trunc),
                            static_cast<double>(n_fin), static_cast<double>(fin_ret), static_cast<double>(fin_len)};
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double tot = block_sum_double(vals[j], s_red);
      if (threadIdx.x == 0) atomicAdd(&io.stats[j], tot);
    }
  }
}

template <class Dyn>
inline void launch_rollout(const typename Dyn::Buffers& b, const RolloutIO& io, int64_t n, const typename Dyn::Consts& k,
                           const RolloutConsts& r, cudaStream_t s) {
  const int grid = grid_for(n, kBlock);
  if (io.rec_obs != nullptr)
    launch_pdl(rollout_f32_kernel<Dyn, true>, grid, kBlock, s, b, io, static_cast<uint32_t>(n), k, r);
  else
    launch_pdl(rollout_f32_kernel<Dyn, false>, grid, kBlock, s, b, io, static_cast<uint32_t>(n), k, r);
}

template <bool IP, int FR>
inline void launch_cartpole_rollout(int ak, float* state, const RolloutIO& io, int64_t n, const CartPoleF32Consts& k,
                                    const RolloutConsts& r, cudaStream_t s) {
  float4* st = reinterpret_cast<float4*>(state);
  switch (ak) {
#define EMEI_AK(A)                                                           \
  case A: {                                                                  \
    typename CartPoleDyn<IP, A, FR>::Buffers b = {st};                       \
    launch_rollout<CartPoleDyn<IP, A, FR>>(b, io, n, k, r, s);               \
  } break;
    EMEI_AK(EMEI_ACTION_DISCRETE_U8)
    EMEI_AK(EMEI_ACTION_DISCRETE_I32)
    EMEI_AK(EMEI_ACTION_DISCRETE_I64)
    EMEI_AK(EMEI_ACTION_CONTINUOUS_F32)
    EMEI_AK(EMEI_ACTION_CONTINUOUS_F64)
#undef EMEI_AK
  }
}

inline void launch_charged_ball_rollout(int ak, uint8_t* on_circle, float* circle, float* free_state, const RolloutIO& io,
                                        int64_t n, const ChargedBallF32Consts& k, const RolloutConsts& r, cudaStream_t s) {
  switch (ak) {
#define EMEI_AK(A)                                                                                                       \
  case A: {                                                                                                              \
    typename ChargedBallDyn<A>::Buffers b = {on_circle, reinterpret_cast<float2*>(circle), reinterpret_cast<float4*>(free_state)}; \
    launch_rollout<ChargedBallDyn<A>>(b, io, n, k, r, s);                                                                \
  } break;
    EMEI_AK(EMEI_ACTION_DISCRETE_U8)
    EMEI_AK(EMEI_ACTION_DISCRETE_I32)
    EMEI_AK(EMEI_ACTION_DISCRETE_I64)
    EMEI_AK(EMEI_ACTION_CONTINUOUS_F32)
    EMEI_AK(EMEI_ACTION_CONTINUOUS_F64)
#undef EMEI_AK
  }
}

}  // namespace emei
