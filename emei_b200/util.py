"""Batched mirror of emei/util.py (a caller of ``env.step``, SURVEY 8b: emei/util.py:16).

The reference's ``random_policy_test(env, is_render, sleep, default_action)`` loops forever: sample an action, step,
accumulate the episode's length and reward, print both when the episode ends, reset.  For a batch of envs that loop is one
fused rollout launch per report (``EmeiEnv.rollout``: policy, step, TimeLimit, per-env reset in-kernel), and what is printed
is the same line with the averages over the episodes that ended inside the report window.
"""
import time

import torch


def random_policy_test(env, is_render=False, sleep=None, default_action=None, report_every=100, max_steps=None, out=print):
    """emei/util.py:5-41 for ``env.num_envs`` envs at once.

    default_action: None = ``env.action_space.sample()`` per env and step (the built-in counter-based random policy);
        otherwise that action for every env and step (emei/util.py:15).
    report_every:   env-steps per report line (one kernel launch each).
    max_steps:      stop after this many steps per env (None = forever, like the reference).
    is_render:      rendering is out of scope for the batched engine; True raises NotImplementedError.
    Returns the list of per-report info dicts (``rollout_info``) when ``max_steps`` is given."""
    if is_render:
        raise NotImplementedError("rendering (pygame / mujoco viewers) is out of scope for the batched engine")
    env.reset()
    reports, done_steps = [], 0
    while max_steps is None or done_steps < max_steps:
        horizon = report_every if max_steps is None else min(report_every, max_steps - done_steps)
        actions = None
        if default_action is not None:
            a = torch.as_tensor(default_action).reshape(-1)[:1]
            actions = a.expand(horizon * env.num_envs).reshape(horizon, env.num_envs).contiguous()
        info = env.rollout_info(env.rollout(horizon, actions=actions)["stats"])
        done_steps += horizon
        if info["total_episode_num"] > 0:  # the reference's line (emei/util.py:25-29), averaged over the finished episodes
            out("episode length: {}\tepisode rewards: {}".format(info["avg_length"], info["avg_reward"]))
        if max_steps is not None:
            reports.append(info)
        if sleep is not None:
            time.sleep(sleep)
    return reports
