"""HalfCheetah scoring: counterpart of ``emei/envs/mujoco/half_cheetah.py`` (reward :59-63, terminal
:65-67, defaults :17-28) on 18-d observations and 6-d actions.  The control cost is summed over the
whole batch (:61) exactly like Hopper's -- see hopper.py in this package."""
from ... import _lib
from .mujoco_env import EmeiMujocoEnv


class HalfCheetahRunningEnv(EmeiMujocoEnv):
    _model = (9, 6, (-1.0, 1.0), [0.0] * 9)  # half_cheetah.xml:89-94
    _family = _lib.HALFCHEETAH

    def __init__(
        self,
        freq_rate: int = 4,
        real_time_scale: float = 0.002,
        integrator="euler",
        forward_reward_weight=1.0,
        ctrl_cost_weight=0.1,
        init_noise_params=0.1,
        obs_noise_params=0.0,
        ctrl_cost_scope: str = "global",
        **kwargs,
    ):
        self._forward_reward_weight = forward_reward_weight
        self._ctrl_cost_weight = ctrl_cost_weight
        self.ctrl_cost_scope = ctrl_cost_scope
        EmeiMujocoEnv.__init__(
            self, observation_dim=18, freq_rate=freq_rate, real_time_scale=real_time_scale, integrator=integrator,
            init_noise_params=init_noise_params, obs_noise_params=obs_noise_params, **kwargs,
        )

    def _scoring_params(self) -> _lib.ScoringParams:
        p = EmeiMujocoEnv._scoring_params(self)
        p.forward_reward_weight = self._forward_reward_weight
        p.ctrl_cost_weight = self._ctrl_cost_weight
        return p
