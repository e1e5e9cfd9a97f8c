"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's collection loop for a batch.

Follows zoo/util.py:33-93 (``rollout``): ``obs = env.reset(); while not done: action = policy | sample;
next_obs, reward, terminated, truncated = env.step(action); record(obs, next_obs, action, reward,
float(done), float(truncated)); obs = next_obs`` with ``done = terminated or truncated``, where
``truncated`` comes from gym's TimeLimit wrapper installed by ``gym.make`` for the registry's
``max_episode_steps`` (register_env.py:14-116; gym 0.26 ``TimeLimit.step``: ``elapsed += 1; truncated =
elapsed >= max_episode_steps``).  Every env of the batch runs that loop independently; an env whose step
was done starts its next episode from a fresh ``reset()`` sample.

The engine draws reset samples and random actions from its counter-based Philox streams
(``oracle/philox.py`` mirrors them bit for bit); numpy's PCG64 stream of the reference cannot be shared.
"""
import numpy as np

from . import philox as P

RESET_STRIDE = 0xD1B54A32D192ED03
PURPOSE_ROLLOUT_ACTION = 4
MASK64 = 0xFFFFFFFFFFFFFFFF


def rollout_seeds(seed: int, epoch: int = 0):
    """(seed_reset, seed_action) exactly as emei_b200.core.EmeiEnv.rollout derives them from reset(seed=);
    ``epoch`` = number of un-seeded reset() calls since that seeded one."""
    seed &= MASK64
    return ((seed * 0x9E3779B97F4A7C15 + 0x5851F42D4C957F2D + epoch * 0xA0761D6478BD642F) & MASK64,
            (seed * 0xD1B54A32D192ED03 + 0x14057B7EF767814F + epoch * 0xE7037ED1A0B428DB) & MASK64)


PURPOSE_ROLLOUT_RESET = 5


def reset_sample_uniform(env_ids, ep_index, seed_reset, low=-0.05, high=0.05, pi_column=-1):
    """float32 reset state of env ``env_ids[j]`` for its ``ep_index[j]``-th in-rollout episode: the lean sampler of
    rollout_f32.cuh -- one Philox block per reset, coordinate c = float32 fma(high - low, u_c, low) with u_c the top
    24 bits of word c (U(low, high) of cartpole.py:131-132,153-156), + pi on the swing-up angle added in float64."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    m = env_ids.shape[0]
    out = np.empty((m, 4), dtype=np.float32)
    lo = np.float32(low)
    span = np.float32(np.float32(high) - lo)
    for j in range(m):
        s = (seed_reset + int(ep_index[j]) * RESET_STRIDE) & MASK64
        w = P.philox4x32_10(s, env_ids[j : j + 1], np.zeros(1), PURPOSE_ROLLOUT_RESET)[:, 0]
        u = (w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)  # exact: 24-bit integer * 2^-24
        v = (np.float64(span) * u.astype(np.float64) + np.float64(lo)).astype(np.float32)  # one rounding = fmaf
        if pi_column >= 0:
            v[pi_column] = np.float32(np.float64(v[pi_column]) + np.pi)
        out[j] = v
    return out


def reset_sample_gaussian(env_ids, ep_index, seed_reset, mean, sigma):
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    out = np.empty((env_ids.shape[0], 4), dtype=np.float32)
    for j in range(env_ids.shape[0]):
        s = (seed_reset + int(ep_index[j]) * RESET_STRIDE) & MASK64
        out[j] = P.init_gaussian(1, 4, mean, sigma, s, env_offset=int(env_ids[j]), dtype=np.float32)[0]
    return out


def reset_sample_charged_ball(env_ids, ep_index, seed_reset, radius=1.0):
    """(on_circle u8[m], circle f32[m,2], free f32[m,4]) of charged_ball.py:84-94 for the in-rollout episodes.
    circle is bit-exact (float32 casts of exact float64 arithmetic); free goes through float32 sin/cos (<= 1 ulp
    from the device's sincosf)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    m = env_ids.shape[0]
    on, circle, free = np.ones(m, np.uint8), np.empty((m, 2), np.float32), np.empty((m, 4), np.float32)
    for j in range(m):
        s = (seed_reset + int(ep_index[j]) * RESET_STRIDE) & MASK64
        _, c, f = P.init_charged_ball(1, radius, s, env_offset=int(env_ids[j]), dtype=np.float32)
        circle[j], free[j] = c[0], f[0]
    return on, circle, free


def random_actions(seed_action, n, t0, horizon, continuous, low=-1.0, high=1.0, env_offset=0):
    """[horizon, n] actions of the engine's uniform random policy (env.action_space.sample() of zoo/util.py:57)."""
    env = np.arange(n, dtype=np.uint64) + np.uint64(env_offset)
    out = np.empty((horizon, n), dtype=np.float32 if continuous else np.uint8)
    for t in range(horizon):
        tg = t0 + t
        if continuous:  # one 32-bit word per step: block tg // 4, word tg % 4, top 24 bits
            w = P.philox4x32_10(seed_action, env, np.full(n, tg >> 2), PURPOSE_ROLLOUT_ACTION)[tg & 3]
            u = (w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
            # float32 fma(high-low, u, low): the product is exact in float64, one rounding at the end
            out[t] = (np.float64(np.float32(high) - np.float32(low)) * u.astype(np.float64) + np.float64(np.float32(low))).astype(np.float32)
        else:  # one bit per step: block tg // 128, word (tg // 32) % 4, bit tg % 32
            w = P.philox4x32_10(seed_action, env, np.full(n, tg >> 7), PURPOSE_ROLLOUT_ACTION)[(tg >> 5) & 3]
            out[t] = ((w >> np.uint32(tg & 31)) & np.uint32(1)).astype(np.uint8)
    return out


def bookkeeping(terminated, max_episode_steps, ep_step, ep_return, rewards):
    """One step of the loop's bookkeeping for a batch (zoo/util.py:58-73 + TimeLimit):
    -> (done, truncated, ep_step', ep_return') BEFORE any reset."""
    ep_step = ep_step + 1
    ep_return = (ep_return.astype(np.float32) + rewards.astype(np.float32)).astype(np.float32)
    truncated = (ep_step >= max_episode_steps) if max_episode_steps > 0 else np.zeros_like(terminated)
    done = terminated | truncated
    return done, truncated, ep_step, ep_return
