"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/emei_oracle_c.c, the plain-C restatement of the reference's
hot path (same role and same rules as oracle/emei_oracle.py: only tests/, smoke() and bench.py's CPU legs may use it).
``make -C oracle`` builds the library; ``available()`` says whether it is there."""
import ctypes
import os
from ctypes import POINTER, Structure, c_double, c_int, c_int64, c_uint8

import numpy as np

from . import emei_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libemei_oracle_c.so")
_lib = None


class _CartPoleParams(Structure):
    _fields_ = [(k, c_double) for k in ("gravity", "mass_pole", "total_mass", "length", "pole_mass_length", "x_threshold", "theta_threshold")]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(LIB_PATH)
        dp, bp = POINTER(c_double), POINTER(c_uint8)
        lib.oc_cartpole_step_f64ref.argtypes = [c_int64, dp, dp, c_double, c_int, POINTER(_CartPoleParams), dp]
        lib.oc_cartpole_reward_terminal.argtypes = [c_int64, dp, c_int, POINTER(_CartPoleParams), dp, bp]
        lib.oc_ip_step.argtypes = [c_int64, dp, dp, c_double, c_int, c_int, c_double, POINTER(_CartPoleParams), dp, dp]
        lib.oc_sumsq.argtypes = [c_int64, dp]
        lib.oc_sumsq.restype = c_double
        lib.oc_hopper_reward_terminal.argtypes = [c_int64, dp, dp] + [c_double] * 4 + [c_int] + [c_double] * 5 + [dp, bp]
        lib.oc_halfcheetah_reward_terminal.argtypes = [c_int64, dp, dp] + [c_double] * 4 + [dp, bp]
        lib.oc_charged_ball_step.argtypes = [c_int64, bp, dp, dp, dp, c_int] + [c_double] * 4 + [c_int]
        lib.oc_charged_ball_step.restype = None
        lib.oc_charged_ball_reward.argtypes = [c_int64, dp, c_double, dp]
        lib.oc_charged_ball_reward.restype = None
        lib.oc_i2p_step.argtypes = [c_int64, dp, dp, c_double, c_int, c_int, dp, dp, dp]
        lib.oc_i2p_step.restype = None
        for f in ("oc_cartpole_step_f64ref", "oc_cartpole_reward_terminal", "oc_ip_step", "oc_hopper_reward_terminal", "oc_halfcheetah_reward_terminal"):
            getattr(lib, f).restype = None
        _lib = lib
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=c_double):
    return a.ctypes.data_as(POINTER(t))


def _cp(p, theta_threshold=0.0):
    return _CartPoleParams(p.gravity, p.mass_pole, p.total_mass, p.length, p.pole_mass_length,
                           float(getattr(p, "x_threshold", 0.0)), float(getattr(p, "theta_threshold_radians", theta_threshold)))


def cartpole_step_f64ref(state, force, dt, freq_rate, p: O.CartPoleParams):
    """base_control.py:72-74,160-164 + cartpole.py:48-60 (oracle/emei_oracle.py cartpole_step_f64ref, libm=True)."""
    s, f = _d(state), _d(np.asarray(force).reshape(-1))
    out = np.empty_like(s)
    cp = _cp(p)
    _load().oc_cartpole_step_f64ref(s.shape[0], _p(s), _p(f), float(dt), int(freq_rate), ctypes.byref(cp), _p(out))
    return out


def cartpole_reward_terminal(kind, obs, p: O.CartPoleParams):
    o = _d(obs)
    r, d = np.empty(o.shape[0]), np.empty(o.shape[0], dtype=np.uint8)
    cp = _cp(p)
    _load().oc_cartpole_reward_terminal(o.shape[0], _p(o), int(kind.endswith("swingup")), ctypes.byref(cp), _p(r), _p(d, c_uint8))
    return r.reshape(-1, 1), d.astype(bool).reshape(-1, 1)


def ip_step(state, ctrl, h, freq_rate, swingup, p: O.InvertedPendulumParams):
    s, c = _d(state), _d(np.asarray(ctrl).reshape(-1))
    out, obs = np.empty_like(s), np.empty_like(s)
    cp = _cp(p)
    _load().oc_ip_step(s.shape[0], _p(s), _p(c), float(h), int(freq_rate), int(bool(swingup)), float(p.gear), ctypes.byref(cp), _p(out), _p(obs))
    return out, obs


def sumsq(action):
    a = _d(action).reshape(-1)
    return float(_load().oc_sumsq(a.shape[0], _p(a)))


def hopper_reward_terminal(obs, pre_obs, action, p: O.HopperParams, sumsq_value=None):
    o, q = _d(obs), _d(pre_obs)
    ss = sumsq(action) if sumsq_value is None else float(sumsq_value)
    r, d = np.empty(o.shape[0]), np.empty(o.shape[0], dtype=np.uint8)
    _load().oc_hopper_reward_terminal(o.shape[0], _p(o), _p(q), ss, p.forward_reward_weight, p.ctrl_cost_weight, p.healthy_reward,
                                      int(p.terminate_when_unhealthy), p.healthy_state_range[0], p.healthy_state_range[1],
                                      p.healthy_z_range[0], p.healthy_z_range[1], p.dt, _p(r), _p(d, c_uint8))
    return r.reshape(-1, 1), d.astype(bool).reshape(-1, 1)


def halfcheetah_reward_terminal(obs, pre_obs, action, p: O.HalfCheetahParams, sumsq_value=None):
    o, q = _d(obs), _d(pre_obs)
    ss = sumsq(action) if sumsq_value is None else float(sumsq_value)
    r, d = np.empty(o.shape[0]), np.empty(o.shape[0], dtype=np.uint8)
    _load().oc_halfcheetah_reward_terminal(o.shape[0], _p(o), _p(q), ss, p.forward_reward_weight, p.ctrl_cost_weight, p.dt, _p(r), _p(d, c_uint8))
    return r.reshape(-1, 1), d.astype(bool).reshape(-1, 1)


def charged_ball_step(on_circle, circle, free, e_force, freq_rate, p: O.ChargedBallParams, f32_force=False):
    """charged_ball.py:25-82 (oracle/emei_oracle.py charged_ball_step, float64, libm=True) -> (on, circle, free)."""
    lib = _load()
    on = np.ascontiguousarray(np.asarray(on_circle).astype(np.uint8))
    ci, fr = _d(circle).copy(), _d(free).copy()
    e = _d(np.asarray(e_force).reshape(-1))
    lib.oc_charged_ball_step(on.shape[0], _p(on, c_uint8), _p(ci), _p(fr), _p(e), int(freq_rate), p.gravity_acc, p.mass_ball, p.radius,
                             p.time_step, int(bool(f32_force)))
    return on.astype(bool), ci, fr


def charged_ball_reward(free, p: O.ChargedBallParams):
    lib = _load()
    f = _d(free)
    r = np.empty(f.shape[0])
    lib.oc_charged_ball_reward(f.shape[0], _p(f), p.radius, _p(r))
    return r.reshape(-1, 1)


def i2p_step(state, ctrl, h, freq_rate, swingup, p: O.I2PParams):
    """oracle/emei_oracle.py i2p_step (float64, libm=True): lagrange_eqs.py:12-60 cartpole(2) + mujoco_env.py:91-97."""
    s, c = _d(state), _d(np.asarray(ctrl).reshape(-1))
    out, obs = np.empty_like(s), np.empty_like(s)
    par = _d([p.mass_cart, p.mass_pole0, p.mass_pole1, p.length0, p.length1, p.gravity, p.gear, p.ctrl_low, p.ctrl_high])
    _load().oc_i2p_step(s.shape[0], _p(s), _p(c), float(h), int(freq_rate), int(bool(swingup)), _p(par), _p(out), _p(obs))
    return out, obs
