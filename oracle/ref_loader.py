"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference (polixir/emei) from /root/reference.

The reference is pure Python/numpy but hard-imports gym, pygame, h5py and mujoco, none of which
are installed here.  This module injects ~100 lines of stub modules into
``sys.modules`` so that ``import emei`` executes the reference's own hot-path code as written.
It exists ONLY so that ``oracle/gen_golden.py`` can generate golden vectors from the executed
reference and so that the CPU tests in this container can pin the numpy restatement
(``oracle/emei_oracle.py``) against it.  ``/root/reference`` does not exist on the GPU box, so
nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this file.

Stubbed third-party behaviour (what the reference relies on):
  * gym.Env.reset(seed=) / np_random  -- gym 0.26 ``seeding.np_random`` = Generator(PCG64(SeedSequence(seed)))
    (call sites: emei/envs/classic_control/base_control.py:44, cartpole.py:132,154)
  * gym.spaces.Discrete / Box         -- only ``contains``, ``sample``, ``shape``, ``dtype``
    (call sites: base_control.py:65-66, cartpole.py:45-46, charged_ball.py:19-20,165-166)
  * gym.envs.mujoco.mujoco_env.MujocoEnv -- only the ``dt`` property (hopper.py:96, half_cheetah.py:60)
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

REFERENCE_ROOT = os.environ.get("EMEI_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "emei"))


class _Space:
    def __init__(self, shape, dtype):
        self.shape = shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))

    def seed(self, seed=None):
        self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        return [seed]


class _Discrete(_Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)

    def contains(self, x):
        if isinstance(x, int):
            v = x
        elif isinstance(x, (np.generic, np.ndarray)) and (np.issubdtype(x.dtype, np.integer) and x.shape == ()):
            v = int(x)
        else:
            return False
        return self.start <= v < self.start + self.n

    def sample(self):
        return int(self.start + self._rng.integers(self.n))


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        low = np.asarray(low, dtype=dtype)
        high = np.asarray(high, dtype=dtype)
        if shape is None:
            shape = low.shape
        super().__init__(tuple(shape), dtype)
        self.low = np.broadcast_to(low, self.shape).copy()
        self.high = np.broadcast_to(high, self.shape).copy()

    def contains(self, x):
        x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype) and x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)
        )

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)


class _Env:
    """Subset of gym 0.26 ``gym.Env``: seeded ``np_random`` and ``reset(seed=)``."""

    metadata = {}
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


_MJ_MODELS = {
    # model_path: (nq, nu, ctrlrange, slider jnt_range, jnt_type (2=slide, 3=hinge), init_qpos)
    # constants read off the reference's MJCF files (emei/envs/mujoco/assets/*.xml)
    "inverted_pendulum.xml": (2, 1, (-3.0, 3.0), (-2.0, 2.0), [2, 3], [0.0, 0.0]),
    "inverted_double_pendulum.xml": (3, 1, (-1.0, 1.0), (-3.0, 3.0), [2, 3, 3], [0.0, 0.0, 0.0]),
    "hopper.xml": (6, 3, (-1.0, 1.0), (-np.inf, np.inf), [2, 2, 3, 3, 3, 3], [0.0, 1.25, 0.0, 0.0, 0.0, 0.0]),
    "half_cheetah.xml": (9, 6, (-1.0, 1.0), (-np.inf, np.inf), [2, 2, 3, 3, 3, 3, 3, 3, 3], [0.0] * 9),
}


class _MujocoEnv(_Env):
    """Stand-in for gym 0.26 ``MujocoEnv`` WITHOUT the MuJoCo library: the constructor records what
    the reference's numpy shell reads (``model.opt.timestep``, ``model.jnt_range``, ``model.jnt_type``,
    ``model.nq/nv``, ``init_qpos/init_qvel``, ``frame_skip``, spaces) from a table of MJCF constants;
    ``dt`` is gym's ``model.opt.timestep * frame_skip`` (used by hopper.py:96, half_cheetah.py:60).
    No dynamics: ``do_simulation`` / ``mj_step`` are not available."""

    def __init__(self, model_path, frame_skip, observation_space, render_mode=None, width=480, height=480,
                 camera_id=None, camera_name=None):
        nq, nu, ctrl, rail, jnt_type, qpos0 = _MJ_MODELS[model_path]
        jnt_range = np.zeros((nq, 2))
        jnt_range[0] = rail
        self.model = SimpleNamespace(
            opt=SimpleNamespace(timestep=self.real_time_scale, integrator=0),
            jnt_range=jnt_range,
            jnt_type=np.array(jnt_type),
            body_quat=np.zeros((nq + 2, 4)),
            nq=nq,
            nv=nq,
            nu=nu,
        )
        self._update_model()  # reference hook (mujoco_env.py:67,111-112), called by _initialize_simulation
        self.frame_skip = frame_skip
        self.observation_space = observation_space
        self.action_space = _Box(low=ctrl[0], high=ctrl[1], shape=(nu,), dtype=np.float32)
        self.init_qpos = np.array(qpos0, dtype=np.float64)
        self.init_qvel = np.zeros(nq, dtype=np.float64)
        self.render_mode = render_mode

    @property
    def dt(self):
        return self.model.opt.timestep * self.frame_skip


class _EzPickle:
    def __init__(self, *a, **k):
        pass


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class DependencyNotInstalled(Exception):
        pass

    spaces = mod("gym.spaces", Discrete=_Discrete, Box=_Box, Space=_Space)
    logger = mod("gym.logger", warn=lambda *a, **k: None)
    error = mod("gym.error", DependencyNotInstalled=DependencyNotInstalled)
    utils = mod("gym.utils", EzPickle=_EzPickle)
    registration = mod(
        "gym.envs.registration",
        registry={},
        register=lambda **k: registration.registry.__setitem__(k["id"], k),
        make=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("gym.make is stubbed")),
        spec=lambda *a, **k: None,
        load_env_plugins=lambda *a, **k: None,
    )
    mj_env = mod("gym.envs.mujoco.mujoco_env", MujocoEnv=_MujocoEnv)
    mj_pkg = mod("gym.envs.mujoco", mujoco_env=mj_env, MujocoEnv=_MujocoEnv)
    envs = mod("gym.envs", registration=registration, mujoco=mj_pkg)
    wrappers = mod("gym.wrappers")
    mod("gym", Env=_Env, spaces=spaces, logger=logger, error=error, utils=utils, envs=envs, wrappers=wrappers)
    gfx = mod("pygame.gfxdraw")
    mod("pygame", gfxdraw=gfx)
    mod("h5py", Dataset=type("Dataset", (), {}), File=None)
    mod("mujoco")


_emei = None


def load_reference():
    """Import and return the reference ``emei`` package (unmodified, from REFERENCE_ROOT)."""
    global _emei
    if _emei is not None:
        return _emei
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("gym", "pygame", "h5py", "mujoco"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", None) is None:
            raise RuntimeError(f"real module {name} already imported; stub loader refuses to shadow it")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import emei  # noqa: E402

    _emei = emei
    return emei


# ---------------------------------------------------------------------------------------------
# Constructors for reference env objects (unmodified classes, constructed around missing deps)
# ---------------------------------------------------------------------------------------------
def make_cartpole(kind: str, freq_rate=1, real_time_scale=0.02, integrator="euler"):
    """kind in {balancing, swingup, continuous_swingup, continuous_balancing}.

    The continuous classes are registered (register_env.py:24-33) but NOT defined in the reference;
    following SURVEY.md 8(a8) they are the discrete class with the continuous-action rule of
    ContinuousChargedBallCenteringEnv (charged_ball.py:163-170): Box(-1,1,(1,),f32), F = mag*a[0].
    """
    load_reference()
    from emei.envs.classic_control.cartpole import CartPoleBalancingEnv, CartPoleSwingUpEnv
    from gym import spaces

    base = CartPoleSwingUpEnv if kind.endswith("swingup") else CartPoleBalancingEnv
    if kind.startswith("continuous"):

        class _Continuous(base):
            def __init__(self, **kw):
                super().__init__(**kw)
                high = np.ones(1, dtype=np.float32)
                self.action_space = spaces.Box(-high, high, dtype=np.float32)

            def _extract_action(self, action):
                return self.force_mag * action[0]

        cls = _Continuous
    else:
        cls = base
    return cls(freq_rate=freq_rate, real_time_scale=real_time_scale, integrator=integrator)


def make_mujoco_shell(name: str, **kw):
    """Construct a reference MuJoCo-family env through its OWN constructor on top of the
    MuJoCo-less ``_MujocoEnv`` stand-in.  name in {hopper, half_cheetah, ip_*, i2p_*}; kw are the
    reference constructor's keyword arguments (freq_rate, real_time_scale, terminate_when_unhealthy...)."""
    load_reference()
    from emei.envs import mujoco as M

    table = {
        "hopper": M.HopperRunningEnv,
        "half_cheetah": M.HalfCheetahRunningEnv,
        "ip_rebound_balancing": M.ReboundInvertedPendulumBalancingEnv,
        "ip_boundary_balancing": M.BoundaryInvertedPendulumBalancingEnv,
        "ip_rebound_swingup": M.ReboundInvertedPendulumSwingUpEnv,
        "ip_boundary_swingup": M.BoundaryInvertedPendulumSwingUpEnv,
        "i2p_rebound_balancing": M.ReboundInvertedDoublePendulumBalancingEnv,
        "i2p_boundary_balancing": M.BoundaryInvertedDoublePendulumBalancingEnv,
        "i2p_rebound_swingup": M.ReboundInvertedDoublePendulumSwingUpEnv,
        "i2p_boundary_swingup": M.BoundaryInvertedDoublePendulumSwingUpEnv,
    }
    return table[name](**kw)


def make_charged_ball(continuous=False, freq_rate=1):
    """The reference class cannot be constructed (charged_ball.py:11-12 passes time_step= to a ctor
    that has no such parameter -> TypeError).  Its physics helpers do run on a hand-built object."""
    load_reference()
    from emei.envs.classic_control.charged_ball import ChargedBallCenteringEnv, ContinuousChargedBallCenteringEnv

    cls = ContinuousChargedBallCenteringEnv if continuous else ChargedBallCenteringEnv
    env = object.__new__(cls)
    env.gravity_acc = 9.8
    env.mass_ball = 1.0
    env.radius = 1.0
    env.charge = 10.0
    env.time_step = 0.02
    env.freq_rate = freq_rate
    env.state = None
    return env


def charged_ball_step(env, action):
    """One env step of the charged ball, driving the reference helpers the way
    BaseControlEnv.step would if the class were constructible: extract action, then freq_rate x
    (update_state(_get_update_info(E))), then obs/reward (charged_ball.py:54-82,96-97,155-160)."""
    e_force = env._extract_action(action)
    for _ in range(env.freq_rate):
        # circle_to_free returns a python LIST (charged_ball.py:28); in free flight update_state does
        # ``list += ndarray`` (:63) which EXTENDS the list and breaks the next unpack.  Holding the
        # free state as an ndarray realises the evidently intended element-wise add without touching
        # the reference's code.
        env.state["free_state"] = np.asarray(env.state["free_state"], dtype=np.float64)
        env.state["circle_state"] = np.asarray(env.state["circle_state"], dtype=np.float64)
        env.update_state(env._get_update_info(e_force))
    return env._get_obs(env.state)
