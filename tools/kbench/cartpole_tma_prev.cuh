// kbench A/B copy (development tool): the TMA step kernel's loop structure BEFORE the trimmed consumer loop of
// the shipped emei_b200/csrc/cartpole_tma.cuh (same integrator), kept to measure what the loop restructuring buys.
#pragma once
#include "../../emei_b200/csrc/cartpole_tma.cuh"
namespace emei {
template <bool IP, int AK, int FR, bool HAS_OBS>
__global__ void __launch_bounds__(kTmaThreads, 1)
    cartpole_step_f32_tma_prev_kernel(const float4* state_in, float4* state_out, float4* obs_out,
                                 const void* __restrict__ action, float* __restrict__ reward, uint8_t* __restrict__ done,
                                 double* stats, uint32_t n, int n_slots, int action_via_tma, const CartPoleF32Consts k) {
  using f32::f2;
  using ActT = typename ActionStorage<AK>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [n_slots] state chunks (8 KB each) | [n_slots] action chunks | full[n_slots] | empty[n_slots] | reduction scratch
  float4* s_state = reinterpret_cast<float4*>(smem_raw);
  ActT* s_act = reinterpret_cast<ActT*>(smem_raw + static_cast<size_t>(n_slots) * kChunk * sizeof(float4));
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_act) + static_cast<size_t>(n_slots) * kChunk * sizeof(ActT));
  uint64_t* empty = full + n_slots;
  const uint32_t tid = threadIdx.x;
  const uint32_t n_chunks = (n + kChunk - 1) / kChunk;
  // this CTA's chunks: blockIdx.x, blockIdx.x + gridDim.x, ...   (local index j)
  const uint32_t my_chunks = blockIdx.x < n_chunks ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const ActT* act = static_cast<const ActT*>(action);

  if (tid == 0) {
    for (int s = 0; s < n_slots; ++s) {
      mbar_init(&full[s], 1);                    // one arrive.expect_tx by the producer; the bytes complete it
      mbar_init(&empty[s], kBlock / 32);         // one arrive per consumer warp of the group that drained the slot
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  pdl_trigger();  // let the next step kernel of the rollout be staged behind this one
  __syncthreads();
  pdl_wait();     // the previous kernel in the stream (the step that wrote state_in) has completed

  float r_acc = 0.0f;
  unsigned d_cnt = 0;
  // ---- producer state (thread 0 only): chunks [0, issued) of this CTA have had their bulk copies issued
  uint32_t issued = 0, p_phase = 0;  // p_phase: parity of the producer slot's CURRENT use
  int p_slot = 0;
  auto issue_until = [&](uint32_t upto) {
    for (; issued < upto; ++issued) {
      const uint32_t c = blockIdx.x + issued * gridDim.x;
      const uint32_t base = c * kChunk;
      const uint32_t cnt = n - base < kChunk ? n - base : kChunk;
      if (issued >= static_cast<uint32_t>(n_slots)) mbar_wait(&empty[p_slot], p_phase ^ 1u);  // previous use drained
      const bool act_tma = action_via_tma && cnt == kChunk;  // partial tail: consumers read their actions directly
      const uint32_t sbytes = cnt * static_cast<uint32_t>(sizeof(float4));
      const uint32_t abytes = act_tma ? kChunk * static_cast<uint32_t>(sizeof(ActT)) : 0u;
      mbar_expect_tx(&full[p_slot], sbytes + abytes);
      tma_load_1d(s_state + static_cast<size_t>(p_slot) * kChunk, state_in + base, sbytes, &full[p_slot]);
      if (act_tma) tma_load_1d(s_act + static_cast<size_t>(p_slot) * kChunk, act + base, abytes, &full[p_slot]);
      if (++p_slot == n_slots) {
        p_slot = 0;
        p_phase ^= 1u;
      }
    }
  };
  if (tid == 0) issue_until(my_chunks < static_cast<uint32_t>(n_slots) ? my_chunks : static_cast<uint32_t>(n_slots));
  {
    // ------------------------------------------------------------------ consumers: group g takes local chunks g, g+4, ...
    const uint32_t g = tid / kBlock, t = tid % kBlock;
    const uint32_t flip = ip_flip(IP, k.variant);
    // 32-bit shared-window addresses, computed once (the generic-pointer forms re-derive the window base per chunk)
    const uint32_t state_u32 = smem_u32(s_state) + t * 16u, full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
    int slot = static_cast<int>(g) % n_slots;
    uint32_t phase = (g / static_cast<uint32_t>(n_slots)) & 1u;
    for (uint32_t j = g; j < my_chunks; j += kTmaGroups) {
      const uint32_t c = blockIdx.x + j * gridDim.x;
      const uint32_t i = c * kChunk + t;  // env A; env B = i + kBlock
      const bool live_a = i < n, live_b = i + kBlock < n;
      const bool act_tma = action_via_tma && (n - c * kChunk >= kChunk);
      if (tid == 0) {  // recycle drained slots: everything up to n_slots chunks ahead of the one consumed now
        const uint32_t ahead = j + static_cast<uint32_t>(n_slots);
        issue_until(my_chunks < ahead ? my_chunks : ahead);
      }
      mbar_wait_u32(full_u32 + static_cast<uint32_t>(slot) * 8u, phase);
      const uint32_t sl = state_u32 + static_cast<uint32_t>(slot) * (kChunk * 16u);
      float4 ya = live_a ? lds128(sl) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 yb = live_b ? lds128(sl + kBlock * 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
      float aa = 0.f, ab = 0.f;
      if (act_tma) {
        const ActT* sa = s_act + static_cast<size_t>(slot) * kChunk;
        aa = static_cast<float>(sa[t]);
        ab = static_cast<float>(sa[t + kBlock]);
      } else {
        if (live_a) aa = static_cast<float>(__ldg(act + i));
        if (live_b) ab = static_cast<float>(__ldg(act + i + kBlock));
      }
      __syncwarp();
      if ((t & 31u) == 0) mbar_arrive_u32(empty_u32 + static_cast<uint32_t>(slot) * 8u);  // this warp's reads of the slot are done
      slot += kTmaGroups;
      while (slot >= n_slots) {
        slot -= n_slots;
        phase ^= 1u;
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
      const float th0a = fabsf(IP ? ya.y : ya.z), th0b = fabsf(IP ? yb.y : yb.z);
      f32::LaneMax<f2> dmax;
      const f2 C = f32::cartpole_integrate<f2, FR>(X, V, TH, W, nf, flip, k.k, k.freq_rate, dmax);
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
      // cos of the reward angle from the integrator (IP: undo the flip of the hanging models)
      float ca, cb;
      f32::f2_unpack(C, ca, cb);
      if constexpr (IP) {
        ca = f32::u2f(f32::f2u(ca) ^ flip);
        cb = f32::u2f(f32::f2u(cb) ^ flip);
      }
      // the integrator's guard (f32math.cuh): otherwise (Inf, absurd angles or rates) redo that env from its
      // stored state with the libm path.  Cold: float32 theta is meaningless there.
      const bool ok_a = th0a <= f32::kSinCosSaneMax && dmax.a <= f32::kDeltaMax;
      const bool ok_b = th0b <= f32::kSinCosSaneMax && dmax.b <= f32::kDeltaMax;
      if (!(ok_a && ok_b)) {
        if (!ok_a && live_a) {
          na = integrate_libm<IP, FR>(state_in[i], fa, flip, k.k, k.freq_rate);
        }
        if (!ok_b && live_b) {
          nb = integrate_libm<IP, FR>(state_in[i + kBlock], fb, flip, k.k, k.freq_rate);
        }
      }
      float rew_a, rew_b;
      bool nd_a, nd_b;
      float4 oa, ob;
      cartpole_outcome<IP>(na, ok_a, ca, k, rew_a, nd_a, oa);
      cartpole_outcome<IP>(nb, ok_b, cb, k, rew_b, nd_b, ob);
      if (live_a) {
        state_out[i] = na;
        if constexpr (HAS_OBS) obs_out[i] = oa;
        reward[i] = rew_a;
        done[i] = nd_a ? 0 : 1;
        r_acc += rew_a;
        d_cnt += nd_a ? 0u : 1u;
      }
      if (live_b) {
        state_out[i + kBlock] = nb;
        if constexpr (HAS_OBS) obs_out[i + kBlock] = ob;
        reward[i + kBlock] = rew_b;
        done[i + kBlock] = nd_b ? 0 : 1;
        r_acc += rew_b;
        d_cnt += nd_b ? 0u : 1u;
      }
    }
  }
  // ---- statistics: warp shuffles -> shared -> one atomic pair per CTA
  if (stats != nullptr) {  // uniform across the grid
    double* s_r = reinterpret_cast<double*>(empty + n_slots);
    unsigned* s_d = reinterpret_cast<unsigned*>(s_r + kTmaThreads / 32);
    const int lane = tid & 31, warp = tid >> 5;
    const double r = warp_sum(static_cast<double>(r_acc));
    const unsigned d = __reduce_add_sync(0xffffffffu, d_cnt);
    if (lane == 0) {
      s_r[warp] = r;
      s_d[warp] = d;
    }
    __syncthreads();
    if (warp == 0) {
      double rr = lane < kTmaThreads / 32 ? s_r[lane] : 0.0;
      unsigned dd = lane < kTmaThreads / 32 ? s_d[lane] : 0u;
      rr = warp_sum(rr);
      dd = __reduce_add_sync(0xffffffffu, dd);
      if (lane == 0) {
        atomicAdd(&stats[0], rr);
        atomicAdd(&stats[1], static_cast<double>(dd));  // exact: counts << 2^53
      }
    }
  }
}

}  // namespace emei
