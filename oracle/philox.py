"""TEST INFRASTRUCTURE ONLY -- numpy mirror of the engine's counter-based sampler
(emei_b200/csrc/common.cuh: Philox4x32-10, Salmon et al. SC'11; kernels.cuh: init_*_kernel).

The reference samples with numpy's PCG64 ``Generator.uniform`` (cartpole.py:131-132,153-156) /
global MT19937 ``randn`` (mujoco_env.py:232-244): third-party bit streams that a device sampler
cannot share, so parity with the REFERENCE is distributional.  This mirror pins the engine's own
stream instead: value(seed, env, column) must be bit-identical on host and device, for any sharding.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
PURPOSE_UNIFORM, PURPOSE_GAUSSIAN, PURPOSE_CHARGED_BALL = 1, 2, 3
PURPOSE_OBS_NOISE = 6


def philox4x32_10(seed: int, env: np.ndarray, block: np.ndarray, purpose: int):
    """-> uint32 [4, N] words for counter (env_lo, env_hi, block, purpose), key (seed_lo, seed_hi)."""
    env = np.asarray(env, dtype=np.uint64)
    c0 = env & MASK
    c1 = env >> np.uint64(32)
    c2 = np.asarray(block, dtype=np.uint64) & MASK
    c3 = np.full_like(c0, purpose, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3]).astype(np.uint32)


def u01(hi, lo):
    a = (hi >> np.uint32(5)).astype(np.float64)
    b = (lo >> np.uint32(6)).astype(np.float64)
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


def uniform_column(seed, env, col, purpose):
    w = philox4x32_10(seed, env, np.full(np.shape(env), col >> 1), purpose)
    return u01(w[2], w[3]) if (col & 1) else u01(w[0], w[1])


def init_uniform(n, dim, low, high, pi_column, seed, env_offset=0, dtype=np.float64):
    env = np.arange(n, dtype=np.uint64) + np.uint64(env_offset)
    out = np.empty((n, dim), dtype=np.float64)
    for c in range(dim):
        v = low + (high - low) * uniform_column(seed, env, c, PURPOSE_UNIFORM)
        if c == pi_column:
            v = v + np.pi
        out[:, c] = v
    return out.astype(dtype)


def init_gaussian(n, dim, mean, sigma, seed, env_offset=0, dtype=np.float64):
    env = np.arange(n, dtype=np.uint64) + np.uint64(env_offset)
    out = np.empty((n, dim), dtype=np.float64)
    for pr in range((dim + 1) // 2):
        w = philox4x32_10(seed, env, np.full(n, pr), PURPOSE_GAUSSIAN)
        u1 = 1.0 - u01(w[0], w[1])
        u2 = u01(w[2], w[3])
        rad = np.sqrt(-2.0 * np.log(u1))
        c0, c1 = 2 * pr, 2 * pr + 1
        out[:, c0] = mean[c0] + sigma[c0] * (rad * np.cos(2 * np.pi * u2))
        if c1 < dim:
            out[:, c1] = mean[c1] + sigma[c1] * (rad * np.sin(2 * np.pi * u2))
    return out.astype(dtype)


def init_charged_ball(n, radius, seed, env_offset=0, dtype=np.float64):
    env = np.arange(n, dtype=np.uint64) + np.uint64(env_offset)
    theta = (-0.5 + uniform_column(seed, env, 0, PURPOSE_CHARGED_BALL)) + np.pi
    omega = -0.5 + uniform_column(seed, env, 1, PURPOSE_CHARGED_BALL)
    T = np.dtype(dtype).type
    th, om = theta.astype(dtype), omega.astype(dtype)
    x, y = np.sin(th) * T(radius), np.cos(th) * T(radius)
    free = np.stack([x, y, om * y, -om * x], axis=1).astype(dtype)
    return np.ones(n, dtype=np.uint8), np.stack([th, om], axis=1), free


def obs_noise(n, dim, sigma, seed, step, freq_rate, env_offset=0):
    """[freq_rate, n, dim] float64: the scaled Gaussian state noise of emei_*_step_noisy (kernels.cuh
    add_state_noise; mujoco_env.py:98-104,197-249): coordinates 2k, 2k+1 of env e at global sub-step
    g = step * freq_rate + s are sigma * (Box-Muller pair of Philox block 4 g + k)."""
    env = np.arange(n, dtype=np.uint64) + np.uint64(env_offset)
    sigma = np.asarray(sigma, dtype=np.float64)
    out = np.empty((freq_rate, n, dim), dtype=np.float64)
    for s in range(freq_rate):
        g = step * freq_rate + s
        for pr in range(dim // 2):
            blk = 4 * g + pr
            w = philox4x32_10(seed, env, np.full(n, blk & 0xFFFFFFFF), PURPOSE_OBS_NOISE | ((blk >> 32) << 8))
            u1 = 1.0 - u01(w[0], w[1])
            u2 = u01(w[2], w[3])
            rad = np.sqrt(-2.0 * np.log(u1))
            out[s, :, 2 * pr] = sigma[2 * pr] * (rad * np.cos(2 * np.pi * u2))
            out[s, :, 2 * pr + 1] = sigma[2 * pr + 1] * (rad * np.sin(2 * np.pi * u2))
    return out
