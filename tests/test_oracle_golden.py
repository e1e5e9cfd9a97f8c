"""Pin the CPU oracle (oracle/emei_oracle.py) against golden vectors produced by EXECUTING the
unmodified reference (oracle/gen_golden.py).  Bit-exact unless stated."""
import warnings

import numpy as np
import pytest

from oracle import emei_oracle as O

warnings.filterwarnings("ignore", category=RuntimeWarning)

CARTPOLE_KINDS = ("balancing", "swingup", "continuous_balancing", "continuous_swingup")


@pytest.mark.parametrize("kind", CARTPOLE_KINDS)
@pytest.mark.parametrize("fr", (1, 4))
def test_cartpole_step_f64ref_bit_exact(golden, kind, fr):
    g = golden("cartpole")
    tag = f"{kind}_fr{fr}"
    p = O.cartpole_params(kind)
    force = O.cartpole_force(g[tag + "_action"], kind.startswith("continuous"), p)
    nxt = O.cartpole_step_f64ref(g[tag + "_state"], force, 0.02, fr, p, libm=True)
    assert np.array_equal(nxt, g[tag + "_next"])
    assert np.array_equal(O.cartpole_reward(kind, nxt), g[tag + "_reward"])
    assert np.array_equal(O.cartpole_terminal(kind, nxt, p), g[tag + "_done"])
    # both outcomes of the terminal test are exercised
    assert 0 < g[tag + "_done"].mean() < 1


@pytest.mark.parametrize("kind", CARTPOLE_KINDS)
@pytest.mark.parametrize("fr", (1, 4))
def test_cartpole_step_f32_within_tolerance(golden, kind, fr):
    """north_star tolerance: 1e-5 rel + 1e-6 abs per teacher-forced step.  It holds for the working
    range |theta| <= 4 pi; beyond that the float32 quantisation of theta itself (ulp(theta) * g/l *
    dt * freq_rate on theta_dot) exceeds 1e-6 abs -- documented in DESIGN.md, checked here with the
    envelope widened by that term."""
    g = golden("cartpole")
    tag = f"{kind}_fr{fr}"
    p = O.cartpole_params(kind)
    st = g[tag + "_state"]
    force = O.cartpole_force(g[tag + "_action"], kind.startswith("continuous"), p)
    n32 = O.cartpole_step_f32(st, force, 0.02, fr, p).astype(np.float64)
    ref = g[tag + "_next"]
    small = np.abs(st[:, 2]) <= 4 * np.pi
    tol = 1e-6 + 1e-5 * np.abs(ref)
    assert np.all(np.abs(n32 - ref)[small] <= tol[small])
    quant = np.spacing(np.abs(st[:, 2]).astype(np.float32)).astype(np.float64)[:, None] * 40.0 * 0.02 * fr
    assert np.all(np.abs(n32 - ref) <= tol + quant)


def test_cartpole_free_running_trajectory(golden):
    g = golden("cartpole")
    p = O.cartpole_params("swingup")
    s = g["traj_swingup_fr4_init"].copy()
    acts, traj = g["traj_swingup_fr4_action"], g["traj_swingup_fr4"]
    for t in range(acts.shape[1]):
        s = O.cartpole_step_f64ref(s, O.cartpole_force(acts[:, t], False, p), 0.02, 4, p, libm=True)
        assert np.array_equal(s, traj[:, t, :4])
        assert np.array_equal(O.cartpole_reward("swingup", s)[:, 0], traj[:, t, 4])
        assert np.array_equal(O.cartpole_terminal("swingup", s, p)[:, 0], traj[:, t, 5].astype(bool))


def test_cartpole_init_state_distribution(golden):
    g = golden("cartpole")
    rng = np.random.default_rng(5)
    for kind in ("swingup", "balancing"):
        ref = g[f"init_{kind}_seed5"]
        mine = O.cartpole_init_state(kind, 4096, rng)
        off = np.array([0, 0, np.pi if kind == "swingup" else 0.0, 0])
        assert np.all(np.abs(ref - off) <= 0.05) and np.all(np.abs(mine - off) <= 0.05)
        assert np.allclose(ref.mean(0), mine.mean(0), atol=3e-3)
        assert np.allclose(ref.std(0), mine.std(0), atol=2e-3)


@pytest.mark.parametrize("T", (1, 0))
def test_hopper_bit_exact(golden, T):
    g = golden("scoring")
    tag = f"hopper_T{T}"
    p = O.HopperParams(terminate_when_unhealthy=bool(T), dt=float(g[tag + "_dt"]))
    obs, pre, act = g[tag + "_obs"], g[tag + "_pre_obs"], g[tag + "_action"]
    assert np.array_equal(O.hopper_reward(obs, pre, act, p), g[tag + "_reward"], equal_nan=True)
    assert np.array_equal(O.hopper_terminal(obs, p), g[tag + "_done"])
    assert np.array_equal(O.hopper_is_healthy(obs, p), g[tag + "_healthy"])
    if T:
        assert not g[tag + "_done"].any()  # hopper.py:105 -- identically False by default
    else:
        assert 0 < g[tag + "_done"].mean() < 1


def test_hopper_known_answers_of_reference_tests(golden):
    """test/test_envs/test_mujoco/test_hopper.py:6-25."""
    g = golden("scoring")
    p = O.HopperParams()
    assert O.hopper_is_healthy(np.ones([128, 12]), p).shape == (128,) and np.all(O.hopper_is_healthy(np.ones([128, 12]), p))
    assert not np.any(O.hopper_is_healthy(np.ones([128, 12]) * 101, p))
    assert np.array_equal(O.hopper_is_healthy(np.ones([128, 12]), p), g["hopper_kat_ones_healthy"])
    assert np.array_equal(O.hopper_is_healthy(np.ones([128, 12]) * 101, p), g["hopper_kat_101_healthy"])
    r = O.hopper_reward(np.ones([128, 12]), np.ones([128, 12]), np.ones([128, 3]), p)
    assert r.shape == (128, 1) and np.array_equal(r, g["hopper_kat_reward"])
    d = O.hopper_terminal(np.ones([128, 12]), p)
    assert d.shape == (128, 1) and np.array_equal(d, g["hopper_kat_done"])


def test_hopper_custom_params(golden):
    g = golden("scoring")
    p = O.HopperParams(
        forward_reward_weight=1.5, ctrl_cost_weight=2e-3, healthy_reward=0.5, terminate_when_unhealthy=False,
        healthy_state_range=(-50.0, 60.0), healthy_z_range=(0.8, 2.0), dt=float(g["hopper_custom_dt"]),
    )
    obs, pre, act = g["hopper_custom_obs"], g["hopper_custom_pre_obs"], g["hopper_custom_action"]
    assert np.array_equal(O.hopper_reward(obs, pre, act, p), g["hopper_custom_reward"])
    assert np.array_equal(O.hopper_terminal(obs, p), g["hopper_custom_done"])


def test_halfcheetah_bit_exact(golden):
    g = golden("scoring")
    p = O.HalfCheetahParams(dt=float(g["halfcheetah_dt"]))
    obs, pre, act = g["halfcheetah_obs"], g["halfcheetah_pre_obs"], g["halfcheetah_action"]
    assert np.array_equal(O.halfcheetah_reward(obs, pre, act, p), g["halfcheetah_reward"], equal_nan=True)
    assert np.array_equal(O.halfcheetah_terminal(obs), g["halfcheetah_done"])
    assert g["halfcheetah_done"].sum() > 0


@pytest.mark.parametrize("kind", ("ip_rebound_balancing", "ip_boundary_balancing", "ip_rebound_swingup", "ip_boundary_swingup"))
def test_ip_reward_terminal(golden, kind):
    g = golden("scoring")
    assert np.array_equal(O.ip_reward(kind, g[kind + "_obs"]), g[kind + "_reward"], equal_nan=True)
    assert np.array_equal(O.ip_terminal(kind, g[kind + "_obs"]), g[kind + "_done"])


@pytest.mark.parametrize("kind", ("i2p_rebound_balancing", "i2p_boundary_balancing", "i2p_rebound_swingup", "i2p_boundary_swingup"))
def test_i2p_reward_terminal(golden, kind):
    g = golden("scoring")
    assert np.array_equal(O.i2p_reward(kind, g[kind + "_obs"]), g[kind + "_reward"], equal_nan=True)
    assert np.array_equal(O.i2p_terminal(kind, g[kind + "_obs"]), g[kind + "_done"])


def test_ip_graph_and_wrap(golden):
    g = golden("scoring")
    for k in (1, 2, 3, 5):
        assert np.array_equal(O.transition_graph_power(O.IP_TRANSITION_GRAPH, 4, 1, k), g[f"ip_graph_k{k}"])
    assert np.array_equal(O.ip_wrap_angle(g["ip_wrap_in"]), g["ip_wrap_out"])


@pytest.mark.parametrize("tag", ("disc_fr1", "disc_fr3", "cont_fr1", "cont_fr3"))
def test_charged_ball_teacher_forced_bit_exact(golden, tag):
    c = golden("charged_ball")
    p = O.ChargedBallParams()
    fr, cont = int(tag[-1]), tag.startswith("cont")
    on, ci, fre, act = c[tag + "_on"], c[tag + "_circle"], c[tag + "_free"], c[tag + "_action"]
    assert 0.05 < 1 - on.mean() < 0.5  # both regimes exercised
    for t in range(act.shape[0]):
        E = O.charged_ball_force(act[t], cont, p)
        o2, c2, f2 = O.charged_ball_step(on[t], ci[t], fre[t], E, fr, p, libm=True, f32_force=cont)
        assert np.array_equal(o2, on[t + 1]) and np.array_equal(f2, fre[t + 1]) and np.array_equal(c2, ci[t + 1])


def test_charged_ball_reward(golden):
    c = golden("charged_ball")
    assert np.array_equal(O.charged_ball_reward(c["reward_free"], O.ChargedBallParams())[:, 0], c["reward"])


def test_ip_dynamics_matches_lagrangian_closed_form():
    """The analytic IP acceleration (parity unpinned vs MuJoCo) equals the single-pole Lagrangian
    solution of auxiliary/lagrange_eqs.py:12-69 (thin rod, inertia 1/3 m l^2 about the COM):
      (M+m) x'' + m l (th'' cos th - th'^2 sin th) = F
      (4/3) m l^2 th'' + m l x'' cos th - m g l sin th = 0."""
    rng = np.random.default_rng(3)
    p = O.InvertedPendulumParams()
    n = 1000
    th, om, v = rng.uniform(-np.pi, np.pi, n), rng.uniform(-8, 8, n), rng.uniform(-5, 5, n)
    F = rng.uniform(-300, 300, n)
    xa, ta = O.cartpole_accel(v, th, om, F, p, np.sin(th), np.cos(th))
    M, m, l, g = p.mass_cart, p.mass_pole, p.length, p.gravity
    r1 = (M + m) * xa + m * l * (ta * np.cos(th) - om**2 * np.sin(th)) - F
    r2 = (4.0 / 3.0) * m * l**2 * ta + m * l * xa * np.cos(th) - m * g * l * np.sin(th)
    assert np.abs(r1).max() < 1e-9 and np.abs(r2).max() < 1e-9


def test_i2p_accelerations_match_both_derivations(golden):
    """oracle.i2p_accel (closed form, LDL^T in the kernel's operation order) against (a) the reference's own
    lagrange_eqs.py cartpole(2) EXECUTED and solved numerically, (b) an independent sympy derivation with the
    complete potential energy (oracle/gen_golden_i2p.py)."""
    g = golden("i2p_dynamics")
    p = O.I2PParams()
    assert np.allclose(g["params"], [p.gravity, p.mass_cart, p.mass_pole0, p.mass_pole1, p.length0, p.length1], rtol=0, atol=0)
    for literal, key in ((False, "acc_physical"), (True, "acc_script")):
        a = O.i2p_accel(g["q"], g["qd"], g["force"], False, p, script_literal=literal)
        assert np.all(np.abs(a - g[key]) <= 1e-11 * np.maximum(1.0, np.abs(g[key])))
    # the script's omission is not a rounding matter
    assert np.abs(g["acc_script"] - g["acc_physical"]).max() > 1.0
    # SwingUp models: pole 0 flipped == theta_0 + pi
    q2 = g["q"].copy()
    q2[:, 1] += np.pi
    assert np.allclose(O.i2p_accel(g["q"], g["qd"], g["force"], True, p), O.i2p_accel(q2, g["qd"], g["force"], False, p), rtol=0, atol=1e-9)


def test_i2p_step_energy_and_obs_quirk():
    p = O.I2PParams()
    rng = np.random.default_rng(3)
    st = rng.uniform(-1, 1, size=(64, 6)) * np.array([1.0, 3.0, 3.0, 1.0, 2.0, 2.0])
    # free motion (no force): total energy drifts only at the forward-Euler rate when h shrinks
    def energy(y):
        x, t0, t1, v, w0, w1 = y.T
        M, m0, m1, l0, l1, g = p.mass_cart, p.mass_pole0, p.mass_pole1, p.length0, p.length1, p.gravity
        v0x, v0y = v + l0 * np.cos(t0) * w0, -l0 * np.sin(t0) * w0
        v1x = v + 2 * l0 * np.cos(t0) * w0 + l1 * np.cos(t0 + t1) * (w0 + w1)
        v1y = -2 * l0 * np.sin(t0) * w0 - l1 * np.sin(t0 + t1) * (w0 + w1)
        T = 0.5 * M * v**2 + 0.5 * m0 * (v0x**2 + v0y**2) + 0.5 * m1 * (v1x**2 + v1y**2) + 0.5 * (m0 * l0**2 / 3) * w0**2 + 0.5 * (m1 * l1**2 / 3) * (w0 + w1) ** 2
        V = m0 * g * l0 * np.cos(t0) + m1 * g * (2 * l0 * np.cos(t0) + l1 * np.cos(t0 + t1))
        return T + V
    e0 = energy(st)
    drift = []
    for fr in (100, 1000):
        y, _ = O.i2p_step(st, np.zeros(64), 0.01 / fr, fr, False, p)
        drift.append(np.abs(energy(y) - e0).max())
    assert drift[1] < drift[0] / 5 and drift[1] < 1e-2  # first-order integrator of an energy-conserving system
    y, obs = O.i2p_step(st, rng.uniform(-2, 2, 64), 0.02, 2, True, p)  # ctrl beyond +-1 is clamped
    y2, _ = O.i2p_step(st, np.clip(rng.uniform(-2, 2, 0), -1, 1) if False else np.zeros(64), 0.02, 2, True, p)
    assert np.array_equal(obs[:, [0, 3, 4, 5]], y[:, [0, 3, 4, 5]])
    assert np.allclose(obs[:, 1:3], (y[:, 1:3] + np.pi) % 2 * np.pi - np.pi, rtol=0, atol=0)


def test_oracle_mirror_symmetry_and_action_hold():
    """Domain properties of the restated dynamics (cartpole.py:48-60, base_control.py:160-164): the map is odd in
    (x, x', theta, theta', F) bit for bit (IEEE arithmetic and libm sin/cos are sign-symmetric), the force is held
    over all sub-steps (freq_rate sub-steps of one call == freq_rate calls with freq_rate = 1), and rewards / terminals
    are even."""
    rng = np.random.default_rng(5)
    n = 20000
    st = rng.uniform(-1, 1, size=(n, 4)) * np.array([4.0, 5.0, 40.0, 8.0])
    act = rng.uniform(-1, 1, size=(n, 1)).astype(np.float32)
    p = O.cartpole_params("continuous_swingup")
    f = O.cartpole_force(act, True, p)
    a = O.cartpole_step_f64ref(st, f, 0.02, 4, p)
    b = O.cartpole_step_f64ref(-st, -f, 0.02, 4, p)
    assert np.array_equal(a, -b)
    assert np.array_equal(O.cartpole_reward("swingup", a), O.cartpole_reward("swingup", b))
    assert np.array_equal(O.cartpole_terminal("swingup", a, p), O.cartpole_terminal("swingup", b, p))
    c = st
    for _ in range(4):
        c = O.cartpole_step_f64ref(c, f, 0.02, 1, p)
    assert np.array_equal(a, c)
    ip = O.InvertedPendulumParams()
    ctrl = rng.uniform(-3, 3, size=n)
    s1, o1 = O.ip_step(st[:, [0, 2, 1, 3]], ctrl, 0.02, 2, True, ip, libm=True)
    s2, o2 = O.ip_step(-st[:, [0, 2, 1, 3]], -ctrl, 0.02, 2, True, ip, libm=True)
    assert np.array_equal(s1, -s2)
    # the observation's (theta + pi) % 2 pi - pi (inverted_pendulum.py:45-49) rounds differently for +theta and -theta
    assert np.allclose(O.ip_reward("ip_boundary_swingup", o1), O.ip_reward("ip_boundary_swingup", o2), rtol=0, atol=1e-14)
