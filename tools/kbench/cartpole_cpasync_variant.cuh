// EXPERIMENT kept for tools/kbench only (not part of the library): the per-thread cp.async (LDGSTS) ring version of
// the packed float32 step kernel.  The shipped kernel is the TMA-staged one (emei_b200/csrc/cartpole_tma.cuh); this
// one measured 10.45-10.7 us per 2^20-env step against 10.76 for TMA at that size and loses at every smaller size
// (profiles/r01_kbench_cartpole_variants.txt).  Findings recorded here: a warp sustains only ~8 outstanding LDGSTS
// (16 per warp ran 40 % slower), and a grid smaller than the kernel's real residency lets PDL-launched successors
// unbalance the SMs.
#pragma once
#include "../../emei_b200/csrc/cartpole_f32.cuh"

namespace emei {

// ---- asynchronous staging (LDGSTS): global -> shared without passing through registers ----------
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

// The step kernel.  Each thread advances TWO envs per iteration (A = block*512 + tid, B = A + 256: two
// coalesced 128-bit streams) with the packed f32x2 arithmetic of f32math.cuh; per-lane work that has no packed
// form (quadrant select / sign of the sincos, MUFU.RCP, reward / done) is done on the two halves.
template <bool IP, int AK, int FR, int MINB, bool HAS_OBS, int S, int DBG = 0>
__global__ void __launch_bounds__(kBlock, MINB)
    cartpole_step_f32_kernel(const float4* state_in, float4* state_out, float4* obs_out,
                             const void* __restrict__ action, float* __restrict__ reward, uint8_t* __restrict__ done,
                             double* stats, uint32_t n, const CartPoleF32Consts k) {
  // 32-bit env index: the launcher splits batches above 2^31 - 2^20 envs (never in practice: that is
  // 32 GiB of float32 state)
  using f32::f2;
  using ActT = typename ActionStorage<AK>::type;
  // actions ride in registers, loaded one iteration ahead: a warp sustains only ~8 outstanding LDGSTS (tools/kbench:
  // 16 per warp ran 40 % slower than 8), and the two 128-bit state rows per stage already use them
  constexpr bool kStageAction = false;
  constexpr uint32_t kPair = 2 * kBlock;
  __shared__ float4 s_state[S][2][kBlock];
  __shared__ ActT s_act[kStageAction ? S : 1][2][kStageAction ? kBlock : 1];
  const uint32_t tid = threadIdx.x;
  const uint32_t stride = gridDim.x * kPair;
  uint32_t i = blockIdx.x * kPair + tid;  // env A of this thread's pair; env B = i + kBlock
  float r_acc = 0.0f;
  unsigned d_cnt = 0;
  const uint32_t flip = ip_flip(IP, k.variant);
  const ActT* act = static_cast<const ActT*>(action);

  // one pair's inputs -> ring slot `slot` (each thread only ever touches its own column of the ring,
  // so completion of its own cp.async groups is all the synchronisation the ring needs)
  auto stage_in = [&](int slot, uint32_t idx) {
    if (DBG != 2) {
      if (idx < n) {
        cp_async<16>(&s_state[slot][0][tid], state_in + idx);
        if constexpr (kStageAction) cp_async<sizeof(ActT)>(&s_act[slot][0][tid], act + idx);
      }
      if (idx + kBlock < n) {
        cp_async<16>(&s_state[slot][1][tid], state_in + idx + kBlock);
        if constexpr (kStageAction) cp_async<sizeof(ActT)>(&s_act[slot][1][tid], act + idx + kBlock);
      }
    }
    cp_async_commit();
  };

  pdl_trigger();  // let the next step kernel of the rollout be staged behind this one
  pdl_wait();     // the previous kernel in the stream (the step that wrote state_in) has completed
#pragma unroll
  for (int d = 0; d < S; ++d) stage_in(d, i + d * stride);  // (i + d*stride cannot wrap: n <= 2^31 - 2^20)
  [[maybe_unused]] float a_reg0 = 0.f, a_reg1 = 0.f;
  if constexpr (!kStageAction) {
    if (i < n) a_reg0 = static_cast<float>(__ldg(act + i));
    if (i + kBlock < n) a_reg1 = static_cast<float>(__ldg(act + i + kBlock));
  }

  int slot = 0;
  while (i < n) {
    const bool live_b = i + kBlock < n;
    cp_async_wait<S - 1>();  // the oldest group (this pair) has landed
    float4 ya = s_state[slot][0][tid];
    float4 yb = live_b ? s_state[slot][1][tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    // DBG != 0 exists only for tools/kbench (1: stores elided, 2: loads elided); the library uses 0
    if constexpr (DBG == 2) {
      ya = make_float4(1e-6f * i, 0.5f, 2e-6f * i, -1.0f);
      yb = make_float4(2e-6f * i, 0.25f, 1e-6f * i, 1.0f);
    }
    float aa, ab;
    if constexpr (kStageAction) {
      aa = static_cast<float>(s_act[slot][0][tid]);
      ab = live_b ? static_cast<float>(s_act[slot][1][tid]) : 0.f;
    } else {
      aa = a_reg0;
      ab = a_reg1;
      const uint32_t i_next = i + stride;
      if (i_next < n) a_reg0 = static_cast<float>(__ldg(act + i_next));
      if (i_next + kBlock < n) a_reg1 = static_cast<float>(__ldg(act + i_next + kBlock));
    }
    const float fa = action_to_f_mt<IP, AK>(aa, k), fb = action_to_f_mt<IP, AK>(ab, k);
    // ---- packed integration: [x, x_dot, theta, theta_dot] (cart-pole) / [x, theta, v, omega] (IP)
    f2 X = f32::f2_pack(ya.x, yb.x), V, TH, W = f32::f2_pack(ya.w, yb.w);
    if constexpr (!IP) {
      V = f32::f2_pack(ya.y, yb.y);
      TH = f32::f2_pack(ya.z, yb.z);
    } else {
      TH = f32::f2_pack(ya.y, yb.y);
      V = f32::f2_pack(ya.z, yb.z);
    }
    const f2 nf = f32::f2_pack(-fa, -fb);
    float tma = fabsf(IP ? ya.y : ya.z), tmb = fabsf(IP ? yb.y : yb.z);
    const int fr = FR > 0 ? FR : k.freq_rate;
#pragma unroll
    for (int sub = 0; sub < fr; ++sub) {
      f32::cartpole_substep2(X, V, TH, W, nf, flip, k.k);
      float ta, tb;
      f32::f2_unpack(TH, ta, tb);
      tma = fmaxf(tma, fabsf(ta));
      tmb = fmaxf(tmb, fabsf(tb));
    }
    stage_in(slot, i + S * stride);  // refill this slot (its previous content is consumed: the results depend on it)
    slot = slot + 1 == S ? 0 : slot + 1;
    float4 na, nb;
    {
      float t0, t1;
      f32::f2_unpack(X, na.x, nb.x);
      f32::f2_unpack(W, na.w, nb.w);
      f32::f2_unpack(V, t0, t1);
      if constexpr (!IP) { na.y = t0; nb.y = t1; } else { na.z = t0; nb.z = t1; }
      f32::f2_unpack(TH, t0, t1);
      if constexpr (!IP) { na.z = t0; nb.z = t1; } else { na.y = t0; nb.y = t1; }
    }
    // the unguarded sincos is valid while |theta| stays below kSinCosSaneMax; otherwise (or NaN) redo that
    // env from its stored state with the libm path.  Cold: float32 theta is meaningless there.
    const bool sane_a = tma <= f32::kSinCosSaneMax, sane_b = tmb <= f32::kSinCosSaneMax;
    bool have_cos = true;
    if (!(sane_a && sane_b)) {
      have_cos = false;
      if (!sane_a) {
        na = state_in[i];
        integrate<IP, FR, true>(na, fa, flip, k);
      }
      if (!sane_b && live_b) {
        nb = state_in[i + kBlock];
        integrate<IP, FR, true>(nb, fb, flip, k);
      }
    }
    // ---- reward angle cosine, packed (cart-pole swing-up: cos(theta); IP: cos(wrapped theta))
    float ca = 0.f, cb = 0.f;
    if (have_cos) {
      if constexpr (!IP) {
        if (k.variant == EMEI_CARTPOLE_SWINGUP) f32::f2_unpack(f32::cos_core(f32::f2_pack(na.z, nb.z)), ca, cb);
      } else {
        f32::f2_unpack(f32::cos_core(f32::f2_pack(wrap_pi_f32(na.y), wrap_pi_f32(nb.y))), ca, cb);
      }
    }
    float rew_a, rew_b;
    bool nd_a, nd_b;
    float4 oa, ob;
    cartpole_outcome<IP>(na, sane_a, have_cos, ca, k, rew_a, nd_a, oa);
    cartpole_outcome<IP>(nb, sane_b, have_cos, cb, k, rew_b, nd_b, ob);
    if (DBG != 1 || na.x == 1234.5f) {
      state_out[i] = na;
      if constexpr (HAS_OBS) obs_out[i] = oa;
      reward[i] = rew_a;
      done[i] = nd_a ? 0 : 1;
    }
    r_acc += rew_a;
    d_cnt += nd_a ? 0u : 1u;
    if (live_b && (DBG != 1 || nb.x == 1234.5f)) {
      state_out[i + kBlock] = nb;
      if constexpr (HAS_OBS) obs_out[i + kBlock] = ob;
      reward[i + kBlock] = rew_b;
      done[i + kBlock] = nd_b ? 0 : 1;
    }
    if (live_b) {
      r_acc += rew_b;
      d_cnt += nd_b ? 0u : 1u;
    }
    i += stride;
  }
  cp_async_wait<0>();
  block_stats_accumulate_counts(stats, static_cast<double>(r_acc), d_cnt);
}

// Launch shape (tools/kbench, B200, 2^20 envs): two envs per thread in packed registers want ~100 registers,
// i.e. 2 resident CTAs of 256 threads per SM.
template <int FR>
struct CartPoleShape {
  static constexpr int kMinBlocks = 2;
  static constexpr int kStages = FR > 0 ? 4 : 2;
};
template <bool IP, int FR>
inline void launch_cartpole_f32(int ak, cudaStream_t s, const float* state_in, float* state_out, float* obs_out,
                                const void* action, int action_bytes, float* reward, uint8_t* done, double* stats,
                                int64_t n, const CartPoleF32Consts& k) {
  constexpr int MB = CartPoleShape<FR>::kMinBlocks, ST = CartPoleShape<FR>::kStages;
  for (int64_t off = 0; off < n; off += kCartPoleMaxLaunch) {
    const int64_t m = n - off < kCartPoleMaxLaunch ? n - off : kCartPoleMaxLaunch;
    const int grid = persistent_grid(m, 2 * kBlock, MB);
    const float4* in4 = reinterpret_cast<const float4*>(state_in) + off;
    float4* out4 = reinterpret_cast<float4*>(state_out) + off;
    float4* obs4 = obs_out ? reinterpret_cast<float4*>(obs_out) + off : nullptr;
    const void* act = static_cast<const char*>(action) + off * action_bytes;
    switch (ak) {
#define EMEI_AK(A)                                                                                                  \
  case A: {                                                                                                         \
    /* 8-byte action encodings: 3 ring stages keep the static shared memory under 48 KB */                         \
    constexpr int STA = (sizeof(typename ActionStorage<A>::type) == 8 && ST > 3) ? 3 : ST;                         \
    if (obs4 != nullptr)                                                                                            \
      launch_pdl(cartpole_step_f32_kernel<IP, A, FR, MB, true, STA>, grid, kBlock, s, in4, out4, obs4, act, reward + off, \
                 done + off, stats, static_cast<uint32_t>(m), k);                                                   \
    else                                                                                                            \
      launch_pdl(cartpole_step_f32_kernel<IP, A, FR, MB, false, STA>, grid, kBlock, s, in4, out4, obs4, act, reward + off, \
                 done + off, stats, static_cast<uint32_t>(m), k);                                                   \
  } break;
      EMEI_AK(EMEI_ACTION_DISCRETE_U8)
      EMEI_AK(EMEI_ACTION_DISCRETE_I32)
      EMEI_AK(EMEI_ACTION_DISCRETE_I64)
      EMEI_AK(EMEI_ACTION_CONTINUOUS_F32)
      EMEI_AK(EMEI_ACTION_CONTINUOUS_F64)
#undef EMEI_AK
    }
  }
}


}  // namespace emei
