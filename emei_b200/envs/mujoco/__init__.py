"""MuJoCo-shell task families of the hot path: the scoring interface of Hopper / HalfCheetah, and the analytic
(closed-form) inverted pendulum and inverted double pendulum, four variants each."""
from . import half_cheetah, hopper, inverted_double_pendulum, inverted_pendulum
from .half_cheetah import HalfCheetahRunningEnv
from .hopper import HopperRunningEnv

_VARIANTS = ("BoundaryInvertedPendulumBalancingEnv", "BoundaryInvertedPendulumSwingUpEnv",
             "ReboundInvertedPendulumBalancingEnv", "ReboundInvertedPendulumSwingUpEnv")
for _name in _VARIANTS:
    globals()[_name] = getattr(inverted_pendulum, _name)
    _double = _name.replace("InvertedPendulum", "InvertedDoublePendulum")
    globals()[_double] = getattr(inverted_double_pendulum, _double)

__all__ = ["HalfCheetahRunningEnv", "HopperRunningEnv", *_VARIANTS,
           *(v.replace("InvertedPendulum", "InvertedDoublePendulum") for v in _VARIANTS)]
