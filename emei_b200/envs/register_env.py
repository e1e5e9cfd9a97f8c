"""Registry: the env ids of ``emei/envs/register_env.py`` (:14-116) -> this package's classes.

gym is not a dependency; ``make(id, **kwargs)`` mirrors ``gym.make`` for these ids (constructor
kwargs pass through, e.g. ``freq_rate``, ``real_time_scale``, ``integrator`` -- zoo/conf/task/*.yaml)
plus the additive ``num_envs`` / ``device`` / ``dtype``.  ``max_episode_steps`` is recorded on the env
(``env.max_episode_steps``); gym's TimeLimit wrapper is third-party behaviour and not re-created.
If real gym is importable, ``register_with_gym()`` registers the same ids there.
"""
import importlib

registry = {}


def register(id: str, entry_point: str, max_episode_steps: int):
    registry[id] = dict(id=id, entry_point=entry_point, max_episode_steps=max_episode_steps)


def spec(id: str):
    return registry[id]


def make(id: str, **kwargs):
    if id not in registry:
        raise KeyError(f"unknown emei env id {id!r}; known: {sorted(registry)}")
    mod_name, cls_name = registry[id]["entry_point"].split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    env = cls(**kwargs)
    env.max_episode_steps = registry[id]["max_episode_steps"]
    env.spec_id = id
    return env


def register_with_gym():
    import gym  # noqa: F401  (optional dependency)
    from gym.envs.registration import register as gym_register

    for r in registry.values():
        gym_register(id=r["id"], entry_point=r["entry_point"], max_episode_steps=r["max_episode_steps"])


_CC = "emei_b200.envs.classic_control"
_MJ = "emei_b200.envs.mujoco"
# Classic (register_env.py:14-43)
register("CartPoleBalancing-v0", f"{_CC}:CartPoleBalancingEnv", 500)
register("CartPoleSwingUp-v0", f"{_CC}:CartPoleSwingUpEnv", 1000)
register("ContinuousCartPoleBalancing-v0", f"{_CC}:ContinuousCartPoleBalancingEnv", 500)
register("ContinuousCartPoleSwingUp-v0", f"{_CC}:ContinuousCartPoleSwingUpEnv", 1000)
register("ChargedBallCentering-v0", f"{_CC}:ChargedBallCenteringEnv", 500)
register("ContinuousChargedBallCentering-v0", f"{_CC}:ContinuousChargedBallCenteringEnv", 1000)
# Mujoco family (register_env.py:47-91)
register("ReboundInvertedPendulumSwingUp-v0", f"{_MJ}:ReboundInvertedPendulumSwingUpEnv", 1000)
register("ReboundInvertedPendulumBalancing-v0", f"{_MJ}:ReboundInvertedPendulumBalancingEnv", 1000)
register("BoundaryInvertedPendulumSwingUp-v0", f"{_MJ}:BoundaryInvertedPendulumSwingUpEnv", 1000)
register("BoundaryInvertedPendulumBalancing-v0", f"{_MJ}:BoundaryInvertedPendulumBalancingEnv", 1000)
register("ReboundInvertedDoublePendulumSwingUp-v0", f"{_MJ}:ReboundInvertedDoublePendulumSwingUpEnv", 1000)
register("ReboundInvertedDoublePendulumBalancing-v0", f"{_MJ}:ReboundInvertedDoublePendulumBalancingEnv", 1000)
register("BoundaryInvertedDoublePendulumSwingUp-v0", f"{_MJ}:BoundaryInvertedDoublePendulumSwingUpEnv", 1000)
register("BoundaryInvertedDoublePendulumBalancing-v0", f"{_MJ}:BoundaryInvertedDoublePendulumBalancingEnv", 1000)
register("HopperRunning-v0", f"{_MJ}:HopperRunningEnv", 1000)
register("HalfCheetahRunning-v0", f"{_MJ}:HalfCheetahRunningEnv", 1000)
