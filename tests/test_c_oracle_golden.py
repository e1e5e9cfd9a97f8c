"""The plain-C restatement of the reference's hot path (oracle/emei_oracle_c.c) against the SAME golden vectors the
numpy restatement is pinned to -- vectors written by executing the unmodified reference (oracle/gen_golden.py) --
bit for bit, and against the numpy restatement on random inputs.  Two independent restatements, one pin."""
import subprocess
import warnings
import os

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import emei_oracle as O

warnings.filterwarnings("ignore", category=RuntimeWarning)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not C.available():
        out = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
    assert C.available()


@pytest.mark.parametrize("kind", ("balancing", "swingup", "continuous_balancing", "continuous_swingup"))
@pytest.mark.parametrize("fr", (1, 4))
def test_c_cartpole_step_bit_exact_vs_executed_reference(golden, kind, fr):
    g = golden("cartpole")
    tag = f"{kind}_fr{fr}"
    p = O.cartpole_params(kind)
    force = O.cartpole_force(g[tag + "_action"], kind.startswith("continuous"), p)
    nxt = C.cartpole_step_f64ref(g[tag + "_state"], force, 0.02, fr, p)
    assert np.array_equal(nxt, g[tag + "_next"])
    r, d = C.cartpole_reward_terminal(kind, nxt, p)
    assert np.array_equal(r, g[tag + "_reward"]) and np.array_equal(d, g[tag + "_done"])


def test_c_cartpole_free_running_trajectory(golden):
    g = golden("cartpole")
    p = O.cartpole_params("swingup")
    s = g["traj_swingup_fr4_init"].copy()
    acts, traj = g["traj_swingup_fr4_action"], g["traj_swingup_fr4"]
    for t in range(acts.shape[1]):
        s = C.cartpole_step_f64ref(s, O.cartpole_force(acts[:, t], False, p), 0.02, 4, p)
        assert np.array_equal(s, traj[:, t, :4])


@pytest.mark.parametrize("T", (0, 1))
def test_c_hopper_bit_exact(golden, T):
    g = golden("scoring")
    tag = f"hopper_T{T}"
    p = O.HopperParams(terminate_when_unhealthy=bool(T), dt=float(g[tag + "_dt"]))
    obs, pre, act = g[tag + "_obs"], g[tag + "_pre_obs"], g[tag + "_action"]
    r, d = C.hopper_reward_terminal(obs, pre, act, p, sumsq_value=float(np.sum(np.square(act))))
    assert np.array_equal(r, g[tag + "_reward"], equal_nan=True) and np.array_equal(d, g[tag + "_done"])
    # its own sequential sum of squares agrees with numpy's pairwise one to rounding
    assert abs(C.sumsq(act) - float(np.sum(np.square(act)))) <= 1e-12 * float(np.sum(np.square(act)))


def test_c_hopper_known_answers_and_custom_params(golden):
    """test/test_envs/test_mujoco/test_hopper.py:6-25 + non-default weights / ranges."""
    g = golden("scoring")
    one = np.ones([128, 12])
    r, d = C.hopper_reward_terminal(one, one, np.ones([128, 3]), O.HopperParams())
    assert np.array_equal(r, g["hopper_kat_reward"]) and np.array_equal(d, g["hopper_kat_done"])
    p = O.HopperParams(forward_reward_weight=1.5, ctrl_cost_weight=2e-3, healthy_reward=0.5, terminate_when_unhealthy=False,
                       healthy_state_range=(-50.0, 60.0), healthy_z_range=(0.8, 2.0), dt=float(g["hopper_custom_dt"]))
    obs, pre, act = g["hopper_custom_obs"], g["hopper_custom_pre_obs"], g["hopper_custom_action"]
    r, d = C.hopper_reward_terminal(obs, pre, act, p, sumsq_value=float(np.sum(np.square(act))))
    assert np.array_equal(r, g["hopper_custom_reward"]) and np.array_equal(d, g["hopper_custom_done"])


def test_c_halfcheetah_bit_exact(golden):
    g = golden("scoring")
    p = O.HalfCheetahParams(dt=float(g["halfcheetah_dt"]))
    obs, pre, act = g["halfcheetah_obs"], g["halfcheetah_pre_obs"], g["halfcheetah_action"]
    r, d = C.halfcheetah_reward_terminal(obs, pre, act, p, sumsq_value=float(np.sum(np.square(act))))
    assert np.array_equal(r, g["halfcheetah_reward"], equal_nan=True) and np.array_equal(d, g["halfcheetah_done"])


def test_c_ip_step_equals_numpy_restatement():
    """The analytic inverted pendulum has no executable reference (MuJoCo's mj_step): the two restatements of the
    reference's closed form must at least agree with each other bit for bit, wrap included."""
    rng = np.random.default_rng(3)
    n = 50000
    st = rng.uniform(-1, 1, size=(n, 4)) * np.array([1.9, 60.0, 5.0, 8.0])
    ctrl = rng.uniform(-3, 3, size=n)
    p = O.InvertedPendulumParams()
    for swingup in (False, True):
        for fr in (1, 3):
            s1, o1 = O.ip_step(st, ctrl, 0.02, fr, swingup, p, libm=True)
            s2, o2 = C.ip_step(st, ctrl, 0.02, fr, swingup, p)
            assert np.array_equal(s1, s2) and np.array_equal(o1, o2)
            assert np.all((o2[:, 1] >= -np.pi) & (o2[:, 1] < np.pi))


@pytest.mark.parametrize("tag", ("disc_fr1", "disc_fr3", "cont_fr1", "cont_fr3"))
def test_c_charged_ball_teacher_forced_bit_exact(golden, tag):
    """charged_ball.py:25-82 driven step by step from the executed reference's own states: regime flags, circle and free
    states bit for bit, both regimes and landings exercised; the continuous variant's float32 force arithmetic
    (NEP 50) included."""
    c = golden("charged_ball")
    p = O.ChargedBallParams()
    fr, cont = int(tag[-1]), tag.startswith("cont")
    on, ci, fre, act = c[tag + "_on"], c[tag + "_circle"], c[tag + "_free"], c[tag + "_action"]
    landed = 0
    for t in range(act.shape[0]):
        o2, c2, f2 = C.charged_ball_step(on[t], ci[t], fre[t], O.charged_ball_force(act[t], cont, p), fr, p, f32_force=cont)
        assert np.array_equal(o2, on[t + 1]) and np.array_equal(f2, fre[t + 1]) and np.array_equal(c2, ci[t + 1])
        landed += int((o2 & ~on[t].astype(bool)).sum())
    assert landed > 0
    assert np.array_equal(C.charged_ball_reward(c["reward_free"], p)[:, 0], c["reward"])


def test_c_i2p_step_equals_numpy_restatement():
    """Same situation as the single pendulum (MuJoCo is the reference's solver): the C and numpy restatements of the
    Lagrangian model -- the numpy one is pinned to both derivations in tests/golden/i2p_dynamics.npz -- agree bit for
    bit, observation quirk included."""
    rng = np.random.default_rng(9)
    n = 30000
    st = rng.uniform(-1, 1, size=(n, 6)) * np.array([2.9, 40.0, 40.0, 4.0, 8.0, 10.0])
    ctrl = rng.uniform(-1.3, 1.3, size=n)
    p = O.I2PParams()
    for swingup in (False, True):
        for fr in (1, 2):
            s1, o1 = O.i2p_step(st, ctrl, 0.02, fr, swingup, p, libm=True)
            s2, o2 = C.i2p_step(st, ctrl, 0.02, fr, swingup, p)
            assert np.array_equal(s1, s2) and np.array_equal(o1, o2)


def test_c_and_numpy_restatements_agree_on_edge_inputs():
    """Empty batches, NaN / Inf states, thresholds hit exactly: the two restatements agree bit for bit (NaN == NaN)."""
    p = O.cartpole_params("swingup")
    assert C.cartpole_step_f64ref(np.zeros((0, 4)), np.zeros(0), 0.02, 4, p).shape == (0, 4)
    r, d = C.cartpole_reward_terminal("swingup", np.zeros((0, 4)), p)
    assert r.shape == (0, 1) and d.shape == (0, 1)
    st = np.array([[np.nan, 0, 0, 0], [0, np.inf, 0.1, 0], [5.0, 0, 0, 0], [-5.0, 0, 0, 0], [4.999999999, 1, np.pi, -3],
                   [0, 0, 1e6, 1e3], [0, 0, -np.inf, 0], [0, 0, 0, 0]], dtype=np.float64)
    f = np.array([10.0, -10.0, 0.0, 3.3, -7.0, 10.0, 1.0, 0.0])
    with np.errstate(all="ignore"):
        a = O.cartpole_step_f64ref(st, f, 0.02, 3, p, libm=True)
        b = C.cartpole_step_f64ref(st, f, 0.02, 3, p)
        assert np.array_equal(a, b, equal_nan=True)
        r, d = C.cartpole_reward_terminal("swingup", st, p)
        assert np.array_equal(d, O.cartpole_terminal("swingup", st, p))  # |x| == 5 and NaN are terminal
        assert np.array_equal(r, O.cartpole_reward("swingup", st), equal_nan=True)
        bal = O.cartpole_params("balancing")
        r, d = C.cartpole_reward_terminal("balancing", st, bal)
        assert np.array_equal(d, O.cartpole_terminal("balancing", st, bal)) and np.all(r == 1.0)
        ip = O.InvertedPendulumParams()
        s1, o1 = O.ip_step(st, np.clip(f, -3, 3), 0.02, 2, True, ip, libm=True)
        s2, o2 = C.ip_step(st, np.clip(f, -3, 3), 0.02, 2, True, ip)
        assert np.array_equal(s1, s2, equal_nan=True) and np.array_equal(o1, o2, equal_nan=True)
    hp = O.HopperParams(terminate_when_unhealthy=False)
    obs = np.ones((4, 12))
    obs[1, 5] = np.nan
    obs[2, 1] = 0.7       # healthy_z is a strict inequality (hopper.py:86-88)
    obs[3, 7] = 100.0     # and so is the state range
    r, d = C.hopper_reward_terminal(obs, obs, np.zeros((4, 3)), hp)
    assert np.array_equal(d, O.hopper_terminal(obs, hp)) and d[:, 0].tolist() == [False, True, True, True]
    assert np.array_equal(r, O.hopper_reward(obs, obs, np.zeros((4, 3)), hp))
