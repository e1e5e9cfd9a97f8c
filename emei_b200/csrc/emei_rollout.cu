// extern "C" entry points of the fused float32 rollouts (rollout_f32.cuh).
#include "rollout_f32.cuh"

namespace {

using namespace emei;

// validation + conversion shared by the two entry points; returns EMEI_OK or an error code
int make_rollout(const emei_rollout_params* r, int max_init_kind, const void* actions, float* rec_observations,
                 float* rec_next_observations, void* rec_actions, float* rec_rewards, uint8_t* rec_dones,
                 uint8_t* rec_timeouts, int32_t* ep_step, float* ep_return, int32_t* ep_index, double* stats,
                 RolloutConsts& rc, RolloutIO& io) {
  EMEI_CHECK_PTR(ep_step);
  EMEI_CHECK_PTR(ep_return);
  EMEI_CHECK_PTR(ep_index);
  if (!r->random_policy) EMEI_CHECK_PTR(actions);
  if (rec_observations != nullptr) {  // records are all-or-nothing
    EMEI_CHECK_PTR(rec_next_observations);
    EMEI_CHECK_PTR(rec_actions);
    EMEI_CHECK_PTR(rec_rewards);
    EMEI_CHECK_PTR(rec_dones);
    EMEI_CHECK_PTR(rec_timeouts);
    EMEI_CHECK_ALIGN16(rec_observations);
    EMEI_CHECK_ALIGN16(rec_next_observations);
  }
  (void)max_init_kind;
  rc.horizon = r->horizon;
  rc.max_episode_steps = r->max_episode_steps;
  rc.auto_reset = r->auto_reset;
  rc.random_policy = r->random_policy;
  rc.init_kind = r->init_kind;
  rc.init_pi_column = r->init_pi_column;
  rc.seed_reset = r->seed_reset;
  rc.seed_action = r->seed_action;
  rc.env_offset = r->env_offset;
  rc.t0 = r->t0;
  rc.init_low = r->init_low;
  rc.init_high = r->init_high;
  for (int j = 0; j < 4; ++j) {
    rc.mean[j] = r->init_mean[j];
    rc.sigma[j] = r->init_sigma[j];
  }
  rc.act_low = static_cast<float>(r->action_low);
  rc.act_high = static_cast<float>(r->action_high);
  io = {ep_step, ep_return, ep_index, actions, reinterpret_cast<float4*>(rec_observations),
        reinterpret_cast<float4*>(rec_next_observations), rec_actions, rec_rewards, rec_dones, rec_timeouts, stats};
  return EMEI_OK;
}

}  // namespace

extern "C" int emei_cartpole_rollout_f32(float* state_io, int32_t* episode_step_io, float* episode_return_io,
                                         int32_t* episode_index_io, const void* actions, float* rec_observations,
                                         float* rec_next_observations, void* rec_actions, float* rec_rewards,
                                         uint8_t* rec_dones, uint8_t* rec_timeouts, double* stats, int64_t n,
                                         const emei_cartpole_params* p, const emei_rollout_params* r,
                                         emei_stream_t stream) {
  if (n < 0 || n > kCartPoleMaxLaunch) return EMEI_ERR_BAD_SIZE;
  EMEI_CHECK_PTR(p);
  EMEI_CHECK_PTR(r);
  if (p->variant < EMEI_CARTPOLE_BALANCING || p->variant > EMEI_IP_BOUNDARY_SWINGUP) return EMEI_ERR_BAD_VARIANT;
  if (p->action_kind < EMEI_ACTION_DISCRETE_U8 || p->action_kind > EMEI_ACTION_CONTINUOUS_F64)
    return EMEI_ERR_BAD_ACTION_KIND;
  if (p->freq_rate < 1 || !(p->dt > 0.0) || r->horizon < 0 || (r->init_kind != 0 && r->init_kind != 1)) return EMEI_ERR_BAD_PARAM;
  if (n == 0 || r->horizon == 0) return EMEI_OK;
  EMEI_CHECK_PTR(state_io);
  EMEI_CHECK_ALIGN16(state_io);
  RolloutConsts rc;
  RolloutIO io;
  const int rcode = make_rollout(r, 1, actions, rec_observations, rec_next_observations, rec_actions, rec_rewards, rec_dones,
                                 rec_timeouts, episode_step_io, episode_return_io, episode_index_io, stats, rc, io);
  if (rcode != EMEI_OK) return rcode;
  const CartPoleF32Consts k = make_cartpole_f32_consts(*p);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool ip = p->variant > EMEI_CARTPOLE_SWINGUP;
  const int ak = p->action_kind;
  if (!ip) {
    if (p->freq_rate == 1) launch_cartpole_rollout<false, 1>(ak, state_io, io, n, k, rc, s);
    else if (p->freq_rate == 4) launch_cartpole_rollout<false, 4>(ak, state_io, io, n, k, rc, s);
    else launch_cartpole_rollout<false, 0>(ak, state_io, io, n, k, rc, s);
  } else {
    if (p->freq_rate == 1) launch_cartpole_rollout<true, 1>(ak, state_io, io, n, k, rc, s);
    else if (p->freq_rate == 4) launch_cartpole_rollout<true, 4>(ak, state_io, io, n, k, rc, s);
    else launch_cartpole_rollout<true, 0>(ak, state_io, io, n, k, rc, s);
  }
  return launch_status();
}

extern "C" int emei_charged_ball_rollout_f32(uint8_t* on_circle_io, float* circle_io, float* free_state_io,
                                             int32_t* episode_step_io, float* episode_return_io, int32_t* episode_index_io,
                                             const void* actions, float* rec_observations, float* rec_next_observations,
                                             void* rec_actions, float* rec_rewards, uint8_t* rec_dones,
                                             uint8_t* rec_timeouts, double* stats, int64_t n,
                                             const emei_charged_ball_params* p, const emei_rollout_params* r,
                                             emei_stream_t stream) {
  if (n < 0 || n > kCartPoleMaxLaunch) return EMEI_ERR_BAD_SIZE;
  EMEI_CHECK_PTR(p);
  EMEI_CHECK_PTR(r);
  if (p->action_kind < EMEI_ACTION_DISCRETE_U8 || p->action_kind > EMEI_ACTION_CONTINUOUS_F64)
    return EMEI_ERR_BAD_ACTION_KIND;
  if (p->freq_rate < 1 || !(p->time_step > 0.0) || !(p->radius > 0.0) || !(p->mass_ball > 0.0) || r->horizon < 0)
    return EMEI_ERR_BAD_PARAM;
  if (n == 0 || r->horizon == 0) return EMEI_OK;
  EMEI_CHECK_PTR(on_circle_io);
  EMEI_CHECK_PTR(circle_io);
  EMEI_CHECK_PTR(free_state_io);
  EMEI_CHECK_ALIGN16(circle_io);
  EMEI_CHECK_ALIGN16(free_state_io);
  RolloutConsts rc;
  RolloutIO io;
  const int rcode = make_rollout(r, 2, actions, rec_observations, rec_next_observations, rec_actions, rec_rewards, rec_dones,
                                 rec_timeouts, episode_step_io, episode_return_io, episode_index_io, stats, rc, io);
  if (rcode != EMEI_OK) return rcode;
  rc.init_kind = 2;  // charged_ball.py:84-94 is the only reset sampler of this family
  launch_charged_ball_rollout(p->action_kind, on_circle_io, circle_io, free_state_io, io, n, make_cb_f32_consts(*p), rc,
                              static_cast<cudaStream_t>(stream));
  return launch_status();
}
