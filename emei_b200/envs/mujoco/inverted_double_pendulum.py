"""Inverted double pendulum: reward / terminal of ``emei/envs/mujoco/inverted_double_pendulum.py``
(:84-90,114-122,150-157,185-196) on 6-d observations ``[x, th1, th2, v, w1, w2]``.

Dynamics are MuJoCo-only in the reference and a SURVEY 8(f) 'next' row here (``step`` raises).
Reference quirk kept visible: the 7x6 causal matrix is stored as ``_causal_graph`` (:42), so the
reference's ``get_transition_graph()`` raises; here the matrix is exposed under both names.
"""
import numpy as np

from ... import _lib
from .mujoco_env import EmeiMujocoEnv


class BaseInvertedDoublePendulumEnv(EmeiMujocoEnv):
    _model = (3, 1, (-1.0, 1.0), [0.0, 0.0, 0.0])  # inverted_double_pendulum.xml:45
    _family = None

    def __init__(self, freq_rate: int = 1, real_time_scale: float = 0.02, integrator="euler",
                 init_noise_params=5e-3, obs_noise_params=0.0, **kwargs):
        EmeiMujocoEnv.__init__(
            self, observation_dim=6, freq_rate=freq_rate, real_time_scale=real_time_scale, integrator=integrator,
            init_noise_params=init_noise_params, obs_noise_params=obs_noise_params, **kwargs,
        )
        self.jnt_range = np.array([[-3.0, 3.0], [-np.inf, np.inf], [-np.inf, np.inf]])  # xml:31
        self._causal_graph = np.array(
            [
                [0, 0, 0, 0, 0, 0],
                [0, 0, 0, 1, 1, 1],
                [0, 0, 0, 1, 1, 1],
                [1, 0, 0, 0, 0, 0],
                [0, 1, 0, 1, 1, 1],
                [0, 0, 1, 1, 1, 1],
                [0, 0, 0, 1, 1, 1],
            ]
        )  # inverted_double_pendulum.py:42-52
        self._transition_graph = self._causal_graph

    def _scoring_params(self) -> _lib.ScoringParams:
        p = EmeiMujocoEnv._scoring_params(self)
        p.x_left, p.x_right = float(self.jnt_range[0][0]), float(self.jnt_range[0][1])
        return p


class ReboundInvertedDoublePendulumBalancingEnv(BaseInvertedDoublePendulumEnv):
    _family = _lib.I2P_REBOUND_BALANCING


class BoundaryInvertedDoublePendulumBalancingEnv(BaseInvertedDoublePendulumEnv):
    _family = _lib.I2P_BOUNDARY_BALANCING


class ReboundInvertedDoublePendulumSwingUpEnv(BaseInvertedDoublePendulumEnv):
    _family = _lib.I2P_REBOUND_SWINGUP


class BoundaryInvertedDoublePendulumSwingUpEnv(BaseInvertedDoublePendulumEnv):
    _family = _lib.I2P_BOUNDARY_SWINGUP
