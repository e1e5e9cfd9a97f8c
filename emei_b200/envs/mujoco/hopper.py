"""Hopper scoring: counterpart of ``emei/envs/mujoco/hopper.py`` (is_healthy :79-93, reward :95-102,
terminal :104-106, constructor defaults :18-36) on 12-d observations and 3-d actions.

Replicated reference behaviour:
  * the control cost is summed over the WHOLE batch: ``np.sum(np.square(action))`` (:98).  When the
    batch is sharded across ranks the local sum is all-reduced so every rank sees the global value.
    ``ctrl_cost_scope="local"`` (additive) keeps the sum per process.
  * ``healthy_angle`` is computed and discarded (third positional argument of np.logical_and is
    ``out=``, :91): only the state and z ranges decide health.
  * with the default ``terminate_when_unhealthy=True`` the terminal is identically False and the
    healthy bonus identically ``healthy_reward`` (:99,105).
"""
from typing import Tuple

from ... import _lib
from ...engine import score
from .mujoco_env import EmeiMujocoEnv


class HopperRunningEnv(EmeiMujocoEnv):
    _model = (6, 3, (-1.0, 1.0), [0.0, 1.25, 0.0, 0.0, 0.0, 0.0])  # hopper.xml:18 (ref=1.25), :36-38
    _family = _lib.HOPPER

    def __init__(
        self,
        freq_rate: int = 4,
        real_time_scale: float = 0.002,
        integrator: str = "rk4",
        init_noise_params=5e-3,
        obs_noise_params=0.0,
        forward_reward_weight: float = 1.0,
        ctrl_cost_weight: float = 1e-3,
        healthy_reward: float = 1.0,
        terminate_when_unhealthy: bool = True,
        healthy_state_range: Tuple[float, float] = (-100.0, 100.0),
        healthy_z_range: Tuple[float, float] = (0.7, float("inf")),
        healthy_angle_range: Tuple[float, float] = (-0.2, 0.2),
        ctrl_cost_scope: str = "global",
        **kwargs,
    ):
        self._forward_reward_weight = forward_reward_weight
        self._ctrl_cost_weight = ctrl_cost_weight
        self._healthy_reward = healthy_reward
        self._terminate_when_unhealthy = terminate_when_unhealthy
        self._healthy_state_range = healthy_state_range
        self._healthy_z_range = healthy_z_range
        self._healthy_angle_range = healthy_angle_range  # unused by the reference's result (hopper.py:91)
        self.ctrl_cost_scope = ctrl_cost_scope
        EmeiMujocoEnv.__init__(
            self, observation_dim=12, freq_rate=freq_rate, real_time_scale=real_time_scale, integrator=integrator,
            init_noise_params=init_noise_params, obs_noise_params=obs_noise_params, **kwargs,
        )

    def _scoring_params(self) -> _lib.ScoringParams:
        p = EmeiMujocoEnv._scoring_params(self)
        p.terminate_when_unhealthy = int(bool(self._terminate_when_unhealthy))
        p.forward_reward_weight = self._forward_reward_weight
        p.ctrl_cost_weight = self._ctrl_cost_weight
        p.healthy_reward = self._healthy_reward
        p.healthy_state_lo, p.healthy_state_hi = self._healthy_state_range
        p.healthy_z_lo, p.healthy_z_hi = self._healthy_z_range
        return p

    def is_healthy(self, next_obs):
        """hopper.py:79-93 -> bool [B] (the terminal of the terminate_when_unhealthy=False path, negated)."""
        p = self._scoring_params()
        p.terminate_when_unhealthy = 0
        _, d, was_np = score(self, p, next_obs, want="terminal")
        h = ~d.reshape(-1)
        return self._ret(h, was_np)
