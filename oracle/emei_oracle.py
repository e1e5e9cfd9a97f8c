"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of polixir/emei's hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file; the product (``emei_b200``) never does.  Every function cites the
reference file:line it restates (paths relative to /root/reference).

Pinning status
--------------
* cart-pole step / reward / terminal / Hopper / HalfCheetah / IP / I2P reward+terminal / charged
  ball helpers: pinned against the EXECUTED reference (oracle/ref_loader.py imports the unmodified
  reference; oracle/gen_golden.py writes tests/golden/*.npz; tests/test_oracle_golden.py checks
  this file against those vectors bit-for-bit).
* analytic inverted-pendulum DYNAMICS: **parity unpinned**.  The reference takes the acceleration
  from MuJoCo's ``mj_step`` (mujoco >= 2.2.0, un-vendored, not installed; call site
  emei/envs/mujoco/mujoco_env.py:93).  What is restated here is the reference's own closed form
  (classic_control/cartpole.py:48-60, algebraically identical to auxiliary/lagrange_eqs.py:12-69)
  with the constants of assets/inverted_pendulum.xml and the forward-Euler position rule of
  mujoco_env.py:91-97.  The numpy shell around it (angle wrap, reward, terminal, graph) IS pinned.
* init-state sampling: distributional only (numpy PCG64 / MT19937 streams are third-party).

Arithmetic notes -- the reference cart-pole step is mixed precision:
``_dsdt`` evaluates in float64 with ``math.sin/cos``, casts the derivative vector to float32
(cartpole.py:60); ``y += derivs(y) * dt`` (base_control.py:164) multiplies float32 by a weak
python float (-> float32 product) and accumulates into the float64 state.
"""
import math
from dataclasses import dataclass, field

import numpy as np

TWO_PI = 2 * np.pi


# =============================================================================================
# cart-pole family  (emei/envs/classic_control/cartpole.py, base_control.py)
# =============================================================================================
@dataclass
class CartPoleParams:
    """Constants of BaseCartPoleEnv.__init__ (cartpole.py:22-31), computed the same way."""

    gravity: float = 9.8
    mass_cart: float = 1.0
    mass_pole: float = 0.1
    length: float = 0.5
    force_mag: float = 10.0
    theta_threshold_radians: float = 12 * 2 * math.pi / 360
    x_threshold: float = 2.4
    total_mass: float = field(init=False)
    pole_mass_length: float = field(init=False)

    def __post_init__(self):
        self.total_mass = self.mass_pole + self.mass_cart  # cartpole.py:25
        self.pole_mass_length = self.mass_pole * self.length  # cartpole.py:51


def cartpole_params(kind: str) -> CartPoleParams:
    p = CartPoleParams()
    if kind.endswith("swingup"):
        p.x_threshold = 5  # cartpole.py:140
    return p


def _libm_map(fn, v):
    """Evaluate a glibc libm function (through python's ``math``) per element of a float64 array."""
    flat = np.asarray(v, dtype=np.float64).ravel()

    def safe(t):
        try:
            return fn(t)
        except ValueError:  # math.* raises where numpy returns nan
            return math.nan

    return np.fromiter((safe(t) for t in flat), np.float64, flat.size).reshape(np.shape(v))


def _sincos64(theta, libm):
    """float64 sin/cos.  libm=True evaluates glibc's sin/cos per element (what ``math.sin`` calls,
    cartpole.py:52-53); libm=False uses numpy's ufunc (may be SVML on AVX512 hosts, <=1 ulp off)."""
    if libm:
        flat = np.asarray(theta, dtype=np.float64).ravel()
        s = np.fromiter((math.sin(v) if math.isfinite(v) else math.nan for v in flat), np.float64, flat.size)
        c = np.fromiter((math.cos(v) if math.isfinite(v) else math.nan for v in flat), np.float64, flat.size)
        return s.reshape(np.shape(theta)), c.reshape(np.shape(theta))
    return np.sin(theta), np.cos(theta)


def cartpole_accel(x_dot, theta, theta_dot, force, p: CartPoleParams, sin_theta, cos_theta):
    """cartpole.py:51-58, same evaluation order (all operands same dtype)."""
    dt_ = sin_theta.dtype.type
    pml = dt_(p.pole_mass_length)
    tm = dt_(p.total_mass)
    temp = (force + pml * (theta_dot * theta_dot) * sin_theta) / tm
    theta_acc = (dt_(p.gravity) * sin_theta - cos_theta * temp) / (
        dt_(p.length) * (dt_(4.0 / 3.0) - dt_(p.mass_pole) * (cos_theta * cos_theta) / tm)
    )
    x_acc = temp - pml * theta_acc * cos_theta / tm
    return x_acc, theta_acc


def cartpole_step_f64ref(state, force, dt, freq_rate, p: CartPoleParams, libm=True):
    """Reference-exact batched cart-pole step: base_control.py:72-74,160-164 + cartpole.py:48-60.

    state [B,4] float64 = [x, x_dot, theta, theta_dot]; force [B] float64 (held over sub-steps:
    the action slot's derivative is 0, cartpole.py:60).  Each of ``freq_rate`` sub-steps advances
    ``dt = real_time_scale`` seconds (base_control.py:73).  Increment = float32(deriv)*float32(dt),
    accumulated in float64.
    """
    y = np.array(state, dtype=np.float64, copy=True)
    force = np.asarray(force, dtype=np.float64).reshape(-1)
    dt32 = np.float32(dt)
    with np.errstate(all="ignore"):
        for _ in range(int(freq_rate)):
            x_dot, theta, theta_dot = y[:, 1], y[:, 2], y[:, 3]
            s, c = _sincos64(theta, libm)
            x_acc, theta_acc = cartpole_accel(x_dot, theta, theta_dot, force, p, s, c)
            derivs = np.stack([x_dot, x_acc, theta_dot, theta_acc], axis=1).astype(np.float32)  # cartpole.py:60
            inc = derivs * dt32  # float32 * weak python float -> float32 (base_control.py:164)
            y += inc.astype(np.float64)
    return y


def cartpole_step_f32(state, force, dt, freq_rate, p: CartPoleParams):
    """All-float32 restatement (the engine's fp32 mode target; tolerance 1e-5 rel + 1e-6 abs vs the
    reference, BASELINE.json north_star)."""
    y = np.array(state, dtype=np.float32, copy=True)
    force = np.asarray(force, dtype=np.float32).reshape(-1)
    dt32 = np.float32(dt)
    with np.errstate(all="ignore"):
        for _ in range(int(freq_rate)):
            x_dot, theta, theta_dot = y[:, 1], y[:, 2], y[:, 3]
            s, c = np.sin(theta), np.cos(theta)
            x_acc, theta_acc = cartpole_accel(x_dot, theta, theta_dot, force, p, s, c)
            y = y + np.stack([x_dot, x_acc, theta_dot, theta_acc], axis=1) * dt32
    return y


def cartpole_force(action, continuous: bool, p: CartPoleParams):
    """_extract_action: discrete ``+mag if a==1 else -mag`` (cartpole.py:121-122,142-143);
    continuous ``mag * a[0]`` (pattern of charged_ball.py:169-170; class absent in the reference)."""
    a = np.asarray(action)
    if continuous:
        a = a.reshape(a.shape[0], -1)[:, 0]
        # reference: python float * np.float32 scalar -> float32 product under NEP 50
        return (np.float32(p.force_mag) * a.astype(np.float32)).astype(np.float64)
    a = a.reshape(-1)
    return np.where(a == 1, p.force_mag, -p.force_mag).astype(np.float64)


def cartpole_reward(kind, obs):
    obs = np.asarray(obs)
    if kind.endswith("swingup"):
        return ((np.cos(obs[:, 2]) + 1) / 2).reshape(-1, 1)  # cartpole.py:149-151
    return np.ones([obs.shape[0], 1])  # cartpole.py:128-129


def cartpole_terminal(kind, obs, p: CartPoleParams):
    obs = np.asarray(obs)
    if kind.endswith("swingup"):
        notdone = np.abs(obs[:, 0]) < p.x_threshold  # cartpole.py:145-147
    else:
        notdone = (np.abs(obs[:, 2]) < p.theta_threshold_radians) & (np.abs(obs[:, 0]) < p.x_threshold)  # :124-126
    return np.logical_not(notdone).reshape(-1, 1)


def cartpole_init_state(kind, batch_size, rng):
    s = rng.uniform(low=-0.05, high=0.05, size=(batch_size, 4))  # cartpole.py:131-132
    if kind.endswith("swingup"):
        s[:, 2] += np.pi  # cartpole.py:153-156
    return s


# =============================================================================================
# analytic inverted pendulum (dynamics: PARITY UNPINNED, see header)
# =============================================================================================
@dataclass
class InvertedPendulumParams:
    """Constants from assets/inverted_pendulum.xml: g (:8), slider range (:14), gear/ctrlrange (:23),
    capsule geoms (:15 cart r=0.1 half-len 0.1; :18 pole r=0.049, fromto length 0.6) at MuJoCo's
    default density 1000 kg/m^3; capsule volume = pi r^2 L + 4/3 pi r^3."""

    gravity: float = 9.81
    gear: float = 100.0
    ctrl_low: float = -3.0
    ctrl_high: float = 3.0
    length: float = 0.3  # half the pole length (COM distance from the hinge)
    x_left: float = -2.0
    x_right: float = 2.0
    mass_cart: float = 1000.0 * (math.pi * 0.1**2 * 0.2 + 4.0 / 3.0 * math.pi * 0.1**3)
    mass_pole: float = 1000.0 * (math.pi * 0.049**2 * 0.6 + 4.0 / 3.0 * math.pi * 0.049**3)
    total_mass: float = field(init=False)
    pole_mass_length: float = field(init=False)

    def __post_init__(self):
        self.total_mass = self.mass_pole + self.mass_cart
        self.pole_mass_length = self.mass_pole * self.length


def ip_wrap_angle(theta):
    """inverted_pendulum.py:45-49: ``(theta + pi) % (2 pi) - pi`` (python/numpy floored modulo)."""
    return (theta + np.pi) % (2 * np.pi) - np.pi


def ip_step(state, ctrl, h, freq_rate, swingup: bool, p: InvertedPendulumParams, dtype=np.float64, libm=False, noise=None):
    """Analytic IP step.  state [B,4] = [x, theta, v, omega] (qpos||qvel, mujoco_env.py:142-144),
    theta UNWRAPPED (the reference wraps only the observation copy, inverted_pendulum.py:45-49).
    Per sub-step (mujoco_env.py:91-97 with integrator="euler"): (q, v) <- (q + v*h, v + a(q,v)*h).
    Acceleration: cartpole.py:51-58 with theta_cartpole = theta (+ pi for SwingUp models, whose
    pole body is flipped by body_quat[2]=[0,0,1,0], inverted_pendulum.py:170-172).  Force =
    gear*ctrl (inverted_pendulum.xml:23).  Returns (new_state, obs) with obs[:,1] wrapped.
    noise: optional [freq_rate, B, 4] float64 -- the already-scaled Gaussian draws that mujoco_env.py:98-104 adds to
    (qpos, qvel) after every sub-step when obs_noise_params != 0 (added in float64, then cast to ``dtype``)."""
    T = np.dtype(dtype).type
    y = np.array(state, dtype=dtype, copy=True)
    force = T(p.gear) * np.asarray(ctrl, dtype=dtype).reshape(-1)
    h = T(h)
    sign = T(-1.0) if swingup else T(1.0)
    with np.errstate(all="ignore"):
        for _ in range(int(freq_rate)):  # `_` indexes the noise of this sub-step
            x, theta, v, omega = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
            if dtype == np.float64:
                s, c = _sincos64(theta, libm)
            else:
                s, c = np.sin(theta), np.cos(theta)
            s, c = sign * s, sign * c
            x_acc, theta_acc = cartpole_accel(v, theta, omega, force, p, s, c)
            y = np.stack([x + v * h, theta + omega * h, v + x_acc * h, omega + theta_acc * h], axis=1)
            if noise is not None:
                y = (y.astype(np.float64) + np.asarray(noise[_], dtype=np.float64)).astype(dtype)
    obs = y.copy()
    obs[:, 1] = ip_wrap_angle(obs[:, 1]).astype(dtype)
    return y, obs


def ip_reward(kind, obs):
    obs = np.asarray(obs)
    if kind.endswith("swingup"):
        y = np.cos(obs[:, 1])
        return ((1 - y) / 2).reshape(-1, 1)  # inverted_pendulum.py:139-142,174-177
    return np.ones([obs.shape[0], 1])  # :73-74,103-104


def ip_terminal(kind, obs, p: InvertedPendulumParams = None):
    p = p or InvertedPendulumParams()
    obs = np.asarray(obs)
    finite = np.isfinite(obs).all(axis=1)
    x = obs[:, 0]
    with np.errstate(invalid="ignore"):
        if kind == "ip_rebound_balancing":
            notdone = (np.cos(obs[:, 1]) >= 0.9) & finite  # :76-79
        elif kind == "ip_boundary_balancing":
            notdone = (np.cos(obs[:, 1]) >= 0) & np.logical_and(p.x_left < x, x < p.x_right) & finite  # :106-111
        elif kind == "ip_rebound_swingup":
            notdone = finite  # :144-146
        elif kind == "ip_boundary_swingup":
            notdone = np.logical_and(p.x_left < x, x < p.x_right) & finite  # :179-183
        else:
            raise KeyError(kind)
    return np.logical_not(notdone).reshape(-1, 1)


IP_TRANSITION_GRAPH = np.array(  # inverted_pendulum.py:39-41 (rows x, theta, v, omega, action)
    [[0, 0, 0, 0], [0, 0, 1, 1], [1, 0, 0, 0], [0, 1, 1, 1], [0, 0, 1, 1]]
)

I2P_CAUSAL_GRAPH = np.array(  # inverted_double_pendulum.py:42-52 (stored as _causal_graph)
    [
        [0, 0, 0, 0, 0, 0],
        [0, 0, 0, 1, 1, 1],
        [0, 0, 0, 1, 1, 1],
        [1, 0, 0, 0, 0, 0],
        [0, 1, 0, 1, 1, 1],
        [0, 0, 1, 1, 1, 1],
        [0, 0, 0, 1, 1, 1],
    ]
)


def transition_graph_power(g, num_obs, num_action, repeat_times):
    """core.py:142-161."""
    g = np.array(g, copy=True)
    assert g.shape == (num_obs + num_action, num_obs)
    if repeat_times == 1:
        return g
    n = num_obs + num_action
    aug = np.zeros([n, n])
    aug[:, :num_obs] = g
    prod = aug.copy()
    acc = np.zeros([n, n])
    for _ in range(repeat_times):
        acc += prod
        prod = np.matmul(prod, aug)
    return (acc > 0).astype(int)[:, :num_obs]


# =============================================================================================
# inverted double pendulum reward / terminal (inverted_double_pendulum.py:84-196)
# =============================================================================================
@dataclass
class I2PParams:
    """Constants from assets/inverted_double_pendulum.xml: gravity (:25), slider range (:31), gear / ctrlrange (:45),
    capsule geoms (:32 cart r=0.1 half-len 0.1; :35,:38 poles r=0.045, fromto length 0.6) at MuJoCo's default
    density 1000 kg/m^3 (capsule volume = pi r^2 L + 4/3 pi r^3).  length = half a pole (hinge -> COM)."""

    gravity: float = 9.81
    gear: float = 500.0
    ctrl_low: float = -1.0
    ctrl_high: float = 1.0
    length0: float = 0.3
    length1: float = 0.3
    x_left: float = -3.0
    x_right: float = 3.0
    mass_cart: float = 1000.0 * (math.pi * 0.1**2 * 0.2 + 4.0 / 3.0 * math.pi * 0.1**3)
    mass_pole0: float = 1000.0 * (math.pi * 0.045**2 * 0.6 + 4.0 / 3.0 * math.pi * 0.045**3)
    mass_pole1: float = 1000.0 * (math.pi * 0.045**2 * 0.6 + 4.0 / 3.0 * math.pi * 0.045**3)


def i2p_accel(q, qd, force, swingup: bool, p: I2PParams, dtype=np.float64, script_literal=False, libm=False):
    """Accelerations [x'', th0'', th1''] of the cart + two poles (relative hinge angles, as MuJoCo's qpos).

    Lagrange equations of ``classic_control/auxiliary/lagrange_eqs.py:12-60`` ``cartpole(2)`` (thin-rod inertia
    1/3 m l^2 about the COM, :37), written as ``A(q) a = b(q, qd, F)`` with the symmetric mass matrix solved by
    LDL^T elimination in a fixed operation order (the CUDA kernel follows the same order):

        A00 = M + m0 + m1          A01 = (m0 + 2 m1) l0 c0 + m1 l1 c01        A02 = m1 l1 c01
        A11 = 4/3 m0 l0^2 + 4 m1 l0^2 + 4 m1 l0 l1 c1 + 4/3 m1 l1^2           A12 = 2 m1 l0 l1 c1 + 4/3 m1 l1^2
        A22 = 4/3 m1 l1^2
        b0 = F + (m0 + 2 m1) l0 s0 w0^2 + m1 l1 s01 (w0 + w1)^2
        b1 = g l0 (m0 + 2 m1) s0 + g l1 m1 s01 + 2 l0 l1 m1 s1 w1 (2 w0 + w1)
        b2 = l1 m1 (g s01 - 2 l0 s1 w0^2)

    ``script_literal=True`` drops the ``2 m1`` of the gravity term of b1, which is what the reference script's
    potential energy yields (it omits the height of pole 1's hinge, lagrange_eqs.py:45); see
    oracle/gen_golden_i2p.py.  SwingUp models flip pole 0 (``body_quat[2] = [0,0,1,0]``,
    inverted_double_pendulum.py:146-148,181-183): theta_0 = 0 hangs down, i.e. theta_0 + pi in the equations."""
    T = np.dtype(dtype).type
    q, qd = np.asarray(q, dtype=dtype), np.asarray(qd, dtype=dtype)
    F = np.asarray(force, dtype=dtype).reshape(-1)
    sign = T(-1.0) if swingup else T(1.0)
    if libm:
        import math as _m

        sin = np.vectorize(_m.sin, otypes=[dtype])
        cos = np.vectorize(_m.cos, otypes=[dtype])
    else:
        sin, cos = np.sin, np.cos
    th0, th1, w0, w1 = q[:, 1], q[:, 2], qd[:, 1], qd[:, 2]
    s0, c0 = sign * sin(th0), sign * cos(th0)
    s1, c1 = sin(th1), cos(th1)
    th01 = th0 + th1
    s01, c01 = sign * sin(th01), sign * cos(th01)
    M, m0, m1, l0, l1, g = (T(v) for v in (p.mass_cart, p.mass_pole0, p.mass_pole1, p.length0, p.length1, p.gravity))
    k_a = (m0 + T(2) * m1) * l0          # (m0 + 2 m1) l0
    k_b = m1 * l1                        # m1 l1
    k_c = T(2) * m1 * l0 * l1            # 2 m1 l0 l1
    k_d = T(4.0 / 3.0) * m1 * l1 * l1    # 4/3 m1 l1^2
    k_e = T(4.0 / 3.0) * m0 * l0 * l0 + T(4) * m1 * l0 * l0 + k_d
    k_g1 = g * l0 * ((m0 + T(2) * m1) if not script_literal else m0)
    a00 = M + m0 + m1
    a01 = k_a * c0 + k_b * c01
    a02 = k_b * c01
    a11 = k_e + T(2) * k_c * c1
    a12 = k_c * c1 + k_d
    a22 = k_d
    ws = w0 + w1
    b0 = F + k_a * s0 * (w0 * w0) + k_b * s01 * (ws * ws)
    b1 = k_g1 * s0 + g * k_b * s01 + k_c * s1 * (w1 * (T(2) * w0 + w1))
    b2 = k_b * (g * s01 - T(2) * l0 * s1 * (w0 * w0))
    # LDL^T
    l10 = a01 / a00
    l20 = a02 / a00
    d1 = a11 - l10 * a01
    t12 = a12 - l10 * a02
    l21 = t12 / d1
    d2 = a22 - l20 * a02 - l21 * t12
    y1 = b1 - l10 * b0
    y2 = b2 - l20 * b0 - l21 * y1
    z2 = y2 / d2
    z1 = y1 / d1 - l21 * z2
    z0 = b0 / a00 - l10 * z1 - l20 * z2
    return np.stack([z0, z1, z2], axis=1)


def i2p_wrap_obs(state):
    """inverted_double_pendulum.py:56-60 ``current_obs``: ``state[1:3] = (state[1:3] + pi) % 2 * pi - pi`` -- the
    precedence makes it ((theta + pi) mod 2) * pi - pi, not a wrap to [-pi, pi) (quirk ledger; replicated)."""
    obs = np.array(state, copy=True)
    obs[:, 1:3] = (obs[:, 1:3] + np.pi) % 2 * np.pi - np.pi
    return obs


def i2p_step(state, ctrl, h, freq_rate, swingup: bool, p: I2PParams, dtype=np.float64, libm=False, noise=None):
    """Analytic I2P step.  state [B,6] = [x, th0, th1, v, w0, w1] (qpos||qvel, mujoco_env.py:142-144), angles
    unwrapped.  Per sub-step (mujoco_env.py:91-97, integrator="euler"): (q, v) <- (q + v h, v + a(q, v) h); force =
    gear * clip(ctrl) (inverted_double_pendulum.xml:45; mj_step clamps ctrl to ctrlrange).  -> (new_state, obs)."""
    T = np.dtype(dtype).type
    y = np.array(state, dtype=dtype, copy=True)
    c = np.clip(np.asarray(ctrl, dtype=dtype).reshape(-1), T(p.ctrl_low), T(p.ctrl_high))
    force = T(p.gear) * c
    h = T(h)
    with np.errstate(all="ignore"):
        for _ in range(int(freq_rate)):
            acc = i2p_accel(y[:, :3], y[:, 3:], force, swingup, p, dtype=dtype, libm=libm)
            q_new = y[:, :3] + y[:, 3:] * h
            v_new = y[:, 3:] + acc * h
            y[:, :3], y[:, 3:] = q_new, v_new
            if noise is not None:  # mujoco_env.py:98-104: [freq_rate, B, 6] scaled Gaussian draws, added in float64
                y = (y.astype(np.float64) + np.asarray(noise[_], dtype=np.float64)).astype(dtype)
        obs = i2p_wrap_obs(y)
    return y, obs


def i2p_reward(kind, obs):
    obs = np.asarray(obs)
    if kind.endswith("balancing"):
        return np.ones([obs.shape[0], 1])  # :84-85,114-115
    y = np.cos(obs[:, 1]) + np.cos(obs[:, 1] + obs[:, 2])
    if kind == "i2p_rebound_swingup":
        return ((2 - y) / 4).reshape(-1, 1)  # :150-153
    omega1, omega2 = obs[:, -2], obs[:, -1]  # :187
    vel_penalty = 5e-3 * omega1**2 + 1e-4 * omega2**2
    return ((2 - y) / 4 - vel_penalty).reshape(-1, 1)  # :185-190


def i2p_terminal(kind, obs, x_left=-3.0, x_right=3.0):
    obs = np.asarray(obs)
    finite = np.isfinite(obs).all(axis=1)
    x = obs[:, 0]
    with np.errstate(invalid="ignore"):
        y = np.cos(obs[:, 1]) + np.cos(obs[:, 1] + obs[:, 2])
        if kind == "i2p_rebound_balancing":
            notdone = (y >= 1.5) & finite  # :87-90
        elif kind == "i2p_boundary_balancing":
            notdone = (y >= 0) & np.logical_and(x_left < x, x < x_right) & finite  # :117-122
        elif kind == "i2p_rebound_swingup":
            notdone = finite  # :155-157
        elif kind == "i2p_boundary_swingup":
            notdone = np.logical_and(x_left < x, x < x_right) & finite  # :192-196
        else:
            raise KeyError(kind)
    return np.logical_not(notdone).reshape(-1, 1)


# =============================================================================================
# Hopper / HalfCheetah reward + terminal (hopper.py:79-106, half_cheetah.py:59-67)
# =============================================================================================
@dataclass
class HopperParams:
    forward_reward_weight: float = 1.0
    ctrl_cost_weight: float = 1e-3
    healthy_reward: float = 1.0
    terminate_when_unhealthy: bool = True
    healthy_state_range: tuple = (-100.0, 100.0)
    healthy_z_range: tuple = (0.7, float("inf"))
    healthy_angle_range: tuple = (-0.2, 0.2)
    dt: float = 0.002 * 4  # gym MujocoEnv.dt = model.opt.timestep * frame_skip


def hopper_is_healthy(obs, p: HopperParams):
    """hopper.py:79-93.  NOTE the angle test is computed and DISCARDED: the third positional
    argument of np.logical_and is ``out=`` (hopper.py:91) -- replicated."""
    obs = np.asarray(obs)
    z = obs[:, 1]
    state = obs[:, 2:]
    lo, hi = p.healthy_state_range
    zlo, zhi = p.healthy_z_range
    with np.errstate(invalid="ignore"):
        healthy_state = np.all(np.logical_and(lo < state, state < hi), axis=1)
        healthy_z = np.logical_and(zlo < z, z < zhi)
    return np.logical_and(healthy_state, healthy_z)


def hopper_reward(obs, pre_obs, action, p: HopperParams, sumsq=None):
    """hopper.py:95-102.  control cost is summed over the WHOLE batch (np.sum(np.square(action)))."""
    obs, pre_obs, action = np.asarray(obs), np.asarray(pre_obs), np.asarray(action)
    x_velocity = (obs[:, 0] - pre_obs[:, 0]) / p.dt
    forward_reward = p.forward_reward_weight * x_velocity
    if sumsq is None:
        sumsq = np.sum(np.square(action))
    control_cost = p.ctrl_cost_weight * sumsq
    healthy_reward = np.logical_or(hopper_is_healthy(obs, p), p.terminate_when_unhealthy) * p.healthy_reward
    return (healthy_reward + forward_reward - control_cost).reshape(-1, 1)


def hopper_terminal(obs, p: HopperParams):
    """hopper.py:104-106 (identically False with the default terminate_when_unhealthy=True)."""
    return (~np.logical_or(hopper_is_healthy(obs, p), p.terminate_when_unhealthy)).reshape(-1, 1)


@dataclass
class HalfCheetahParams:
    forward_reward_weight: float = 1.0
    ctrl_cost_weight: float = 0.1
    dt: float = 0.002 * 4


def halfcheetah_reward(obs, pre_obs, action, p: HalfCheetahParams, sumsq=None):
    """half_cheetah.py:59-63."""
    obs, pre_obs, action = np.asarray(obs), np.asarray(pre_obs), np.asarray(action)
    forward_reward = p.forward_reward_weight * (obs[:, 0] - pre_obs[:, 0]) / p.dt
    if sumsq is None:
        sumsq = np.sum(np.square(action))
    control_cost = p.ctrl_cost_weight * sumsq
    return (forward_reward - control_cost).reshape(-1, 1)


def halfcheetah_terminal(obs):
    """half_cheetah.py:65-67."""
    return np.logical_not(np.isfinite(np.asarray(obs)).all(axis=1)).reshape(-1, 1)


def mujoco_init_state(init_qpos, init_qvel, batch_size, noise, rng):
    """INTENDED semantics of mujoco_env.py:137-140,197-249 for slide/hinge joints: every joint
    coordinate gets an independent N(0, sigma) sample per batch row.  (The reference slices ROWS
    instead of columns at :243-244 -- ``noisy_pos[cur_pos_idx:cur_pos_idx+1]`` -- so it raises
    ValueError for batch_size>1 and for batch_size=1 adds ONE shared sample to every coordinate;
    documented in DESIGN.md, not replicated.)  noise: scalar or (sigma_pos, sigma_vel)."""
    sp, sv = (noise if isinstance(noise, tuple) else (noise, noise))
    nq, nv = len(init_qpos), len(init_qvel)
    pos = np.tile(np.asarray(init_qpos, dtype=np.float64)[None], [batch_size, 1]) + rng.standard_normal((batch_size, nq)) * sp
    vel = np.tile(np.asarray(init_qvel, dtype=np.float64)[None], [batch_size, 1]) + rng.standard_normal((batch_size, nv)) * sv
    return pos, vel


# =============================================================================================
# charged ball (emei/envs/classic_control/charged_ball.py)
# =============================================================================================
@dataclass
class ChargedBallParams:
    gravity_acc: float = 9.8  # charged_ball.py:13-17
    mass_ball: float = 1.0
    radius: float = 1.0
    charge: float = 10.0
    time_step: float = 0.02


def _cb_get_angle(x, y, radius, T, libm=False):
    """charged_ball.py:30-36 (vectorised).  Python ``%`` is the floored modulo (= np.mod)."""
    scale = np.sqrt(x * x + y * y)
    arg = x / (scale * T(radius) + T(1e-8))
    a = _libm_map(math.asin, arg) if libm else np.arcsin(arg)
    angle = np.where(y > 0, a, T(np.pi) - a)
    return np.mod(angle, T(2 * np.pi))


def charged_ball_step(on_circle, circle, free, e_force, freq_rate, p: ChargedBallParams, dtype=np.float64, libm=False,
                      f32_force=False):
    """Batched charged-ball step = freq_rate x (update_state(_get_update_info(E))),
    charged_ball.py:54-82 with helpers :25-52.  on_circle bool[B]; circle [B,2]=[theta, omega];
    free [B,4]=[x, y, vx, vy]; e_force [B].  Sub-step h = time_step / freq_rate (:58,63).
    libm=True (float64 only) evaluates sin/cos/asin with glibc per element like ``math.*`` does.

    f32_force=True (float64 only) replicates what the reference does for the CONTINUOUS variant when
    executed under numpy >= 2 (NEP 50): ``self.charge * action[0]`` (:170) is a float32 scalar, and
    every ``python_float (op) np.float32`` in _get_update_info (:75-77, 80) is then evaluated in
    float32: cos*E, (sin*gravity + cos*E), .../(m*r), sin*E and E/m."""
    T = np.dtype(dtype).type
    libm = bool(libm) and np.dtype(dtype) == np.float64

    def sincos(v):
        return _sincos64(v, True) if libm else (np.sin(v), np.cos(v))

    on = np.array(on_circle, dtype=bool, copy=True)
    circle = np.array(circle, dtype=dtype, copy=True)
    free = np.array(free, dtype=dtype, copy=True)
    E = np.asarray(e_force, dtype=dtype).reshape(-1)
    f32_force = bool(f32_force) and np.dtype(dtype) == np.float64
    E32 = E.astype(np.float32)
    F = np.float32
    h = T(p.time_step / freq_rate)
    m, g, r = T(p.mass_ball), T(p.gravity_acc), T(p.radius)
    with np.errstate(all="ignore"):
        for _ in range(int(freq_rate)):
            theta, omega = circle[:, 0], circle[:, 1]
            x, y, vx, vy = free[:, 0], free[:, 1], free[:, 2], free[:, 3]
            # ---- on-circle branch (:72-78, 56-61, 25-28)
            s, c = sincos(theta)
            centrifugal = m * (omega * omega) * r
            gravity = m * g
            if f32_force:
                theta_acc = (((s * gravity).astype(F) + c.astype(F) * E32) / F(m * r)).astype(np.float64)
                flag = centrifugal + (s.astype(F) * E32).astype(np.float64) < c * gravity
                acc_x = (E32 / F(m)).astype(np.float64)
            else:
                theta_acc = (s * gravity + c * E) / (m * r)
                flag = centrifugal + s * E < c * gravity
                acc_x = E / m
            theta_n = theta + omega * h
            omega_n = omega + theta_acc * h
            sn, cn = sincos(theta_n)
            xc, yc = sn * r, cn * r
            free_c = np.stack([xc, yc, omega_n * yc, -omega_n * xc], axis=1)
            # ---- free-flight branch (:79-82, 62-66, 44-52)
            free_f = np.stack([x + vx * h, y + vy * h, vx + acc_x * h, vy + (-g) * h], axis=1)
            xf, yf, vxf, vyf = free_f[:, 0], free_f[:, 1], free_f[:, 2], free_f[:, 3]
            land = xf * xf + yf * yf > r * r + T(0.001)
            th_land = _cb_get_angle(xf, yf, p.radius, T, libm)
            v_angle = _cb_get_angle(vxf, vyf, p.radius, T, libm)
            d = v_angle - th_land
            greater = np.where(np.abs(d) < T(np.pi), v_angle > th_land, v_angle < th_land)  # :38-42
            speed = np.sqrt(vxf * vxf + vyf * vyf) / r
            om_land = np.where(greater, speed, -speed)
            # ---- select
            new_circle = np.where(
                on[:, None],
                np.stack([theta_n, omega_n], axis=1),
                np.where(land[:, None], np.stack([th_land, om_land], axis=1), circle),
            )
            new_free = np.where(on[:, None], free_c, free_f)
            new_on = np.where(on, ~flag, land)
            circle, free, on = new_circle.astype(dtype), new_free.astype(dtype), new_on
    return on, circle, free


def charged_ball_force(action, continuous, p: ChargedBallParams):
    a = np.asarray(action)
    if continuous:
        a = a.reshape(a.shape[0], -1)[:, 0]
        return (np.float32(p.charge) * a.astype(np.float32)).astype(np.float64)  # charged_ball.py:169-170
    a = a.reshape(-1)
    return np.where(a == 1, p.charge, -p.charge).astype(np.float64)  # :155-156


def charged_ball_reward(free, p: ChargedBallParams):
    """charged_ball.py:158-160 / 172-174 (scalar in the reference; batched here)."""
    free = np.asarray(free)
    T = free.dtype.type
    x, y = free[:, 0], free[:, 1]
    return (T(1) - np.sqrt(x * x + y * y) / T(p.radius)).reshape(-1, 1)


def charged_ball_terminal(free):
    """charged_ball.py:110-111: always False."""
    return np.zeros([np.asarray(free).shape[0], 1], dtype=bool)


def charged_ball_init_state(batch_size, rng, p: ChargedBallParams):
    """charged_ball.py:84-94: [theta, omega] = U(-.5,.5,2) + [pi, 0]; on_circle; free via circle_to_free."""
    circle = rng.uniform(low=-0.5, high=0.5, size=(batch_size, 2)) + np.array([np.pi, 0.0])
    x, y = np.sin(circle[:, 0]) * p.radius, np.cos(circle[:, 0]) * p.radius
    free = np.stack([x, y, circle[:, 1] * y, -circle[:, 1] * x], axis=1)
    return np.ones(batch_size, dtype=bool), circle, free
