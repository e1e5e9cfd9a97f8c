#!/usr/bin/env python
"""Times the UNMODIFIED reference's own scalar env loop ("reference-literal" CPU baseline):

    obs, _ = env.reset(seed=...); for t in range(T): obs, r, term, trunc, info = env.step(a)   # base_control.py:61-83

for ContinuousCartPoleSwingUp (freq_rate = 4, BASELINE configs[1]) and BoundaryInvertedPendulumSwingUp's cart-pole
counterpart at freq_rate = 1 (configs[0]; the reference's own IP step is MuJoCo, absent here), on ONE core and on all
cores (one independent env per process -- the reference has no batching).  Runs only in the BUILD container
(it imports /root/reference through oracle/ref_loader.py); the GPU box has no reference, so bench.py carries the
result as a stated constant: `profiles/cpu_literal.json` -> bench line key `cpu_baseline_literal`.

    python scripts/time_reference_literal.py            # writes profiles/cpu_literal.json
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _loop(args):
    kind, freq_rate, steps, seed = args
    from oracle import ref_loader as RL

    env = RL.make_cartpole(kind, freq_rate=freq_rate)
    env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    cont = kind.startswith("continuous")
    acts = rng.uniform(-1, 1, size=(steps, 1)).astype(np.float32) if cont else rng.integers(0, 2, size=steps)
    n = 0
    t0 = time.perf_counter()
    for t in range(steps):
        _, _, term, _, _ = env.step(acts[t] if cont else int(acts[t]))
        n += 1
        if term:
            env.reset()
    return n, time.perf_counter() - t0


def rate(kind, freq_rate, steps, procs):
    if procs == 1:
        _loop((kind, freq_rate, 2000, 0))  # warm imports / caches
        n, wall = _loop((kind, freq_rate, steps, 1))
        return n / wall
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_loop, [(kind, freq_rate, 500, 100 + i) for i in range(procs)])
        t0 = time.perf_counter()
        out = pool.map(_loop, [(kind, freq_rate, steps, 1 + i) for i in range(procs)])
        wall = time.perf_counter() - t0
    return sum(n for n, _ in out) / wall


def main():
    from oracle import ref_loader as RL

    if not RL.reference_available():
        raise SystemExit("the reference tree is not present: this script runs in the build container only")
    cores = len(os.sched_getaffinity(0))
    cpu = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                cpu = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    out = {
        "what": "the unmodified reference's scalar env.step loop (base_control.py:61-83), executed from /root/reference "
                "under the stub modules of oracle/ref_loader.py; one env per process (the reference has no batching)",
        "script": "scripts/time_reference_literal.py",
        "where": "build container (no GPU); NOT the GPU box -- carried into the bench line as a stated constant",
        "cpu": cpu, "cores_available": cores, "python": platform.python_version(), "numpy": np.__version__,
        "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        "unit": "env-steps/s",
        "workloads": {},
    }
    for key, kind, fr in (("c2", "continuous_swingup", 4), ("c1_cartpole_counterpart", "swingup", 1)):
        r1 = rate(kind, fr, 30000, 1)
        rn = rate(kind, fr, 15000, cores)
        out["workloads"][key] = {"env": kind, "freq_rate": fr, "one_core": r1, "all_cores": rn, "processes": cores}
        print(key, f"1 core {r1:.0f} env-steps/s, {cores} processes {rn:.0f} env-steps/s")
    # the reference's own (already vectorised) get_batch_reward + get_batch_terminal on float64 numpy arrays, one process
    # (hopper.py:79-106, half_cheetah.py:59-67 executed from /root/reference through the MuJoCo-less shell of ref_loader)
    out["unit_scoring"] = "transitions/s"
    for key, name, d, da, kw in (("c3_hopper", "hopper", 12, 3, dict(terminate_when_unhealthy=False)), ("c3_halfcheetah", "half_cheetah", 18, 6, {})):
        env = RL.make_mujoco_shell(name, **kw)
        rng = np.random.default_rng(3)
        n = 1 << 20
        obs = rng.standard_normal((n, d))
        pre = obs + 0.01 * rng.standard_normal((n, d))
        act = rng.uniform(-1, 1, size=(n, da))
        env.get_batch_reward(obs, pre, act), env.get_batch_terminal(obs, pre, act)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            env.get_batch_reward(obs, pre, act)
            env.get_batch_terminal(obs, pre, act)
        r1 = n * reps / (time.perf_counter() - t0)
        out["workloads"][key] = {"env": name, "one_core": r1, "all_cores": None, "processes": 1, "batch": n,
                                 "note": "get_batch_reward + get_batch_terminal of the unmodified reference class, float64 numpy, 2^20 transitions per call"}
        print(key, f"1 process {r1:.0f} transitions/s")
    path = os.path.join(ROOT, "profiles", "cpu_literal.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
