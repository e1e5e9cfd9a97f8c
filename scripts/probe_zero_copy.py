"""Zero-copy probe (development): a C2 host step whose kernel stores obs / reward / done straight into PINNED HOST memory
(UVA: a page-locked host pointer is a device pointer) instead of device buffers + three D2H copies per range.

Modes:  A  env.step_host as shipped (copy engines, CUDA graph)
        B  H2D(actions) + ONE kernel whose outputs are pinned host tensors
        C  the same in R ranges on R streams (range k+1's upload overlaps range k's stores)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import emei_b200 as E  # noqa: E402

n = 1 << 20
rng = np.random.default_rng(0)
st = (rng.uniform(-1, 1, size=(n, 4)) * np.array([4.0, 5.0, np.pi, 8.0])).astype(np.float32)
act = torch.as_tensor(rng.uniform(-1, 1, size=n).astype(np.float32)).pin_memory()


def make():
    env = E.make("ContinuousCartPoleSwingUp-v0", freq_rate=4, num_envs=n, dtype=torch.float32)
    env.state = st
    return env


def timeit(fn, reps=40, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


env = make()
dtA = timeit(lambda: env.step_host(act))
refA = [np.array(x) for x in env.step_host(act)[:3]]
print(f"A step_host (shipped)          : {dtA*1e3:.3f} ms/step  {n/dtA/1e9:.2f} G env-steps/s", flush=True)

obs_h = torch.empty((n, 4), dtype=torch.float32).pin_memory()
rew_h = torch.empty((n, 1), dtype=torch.float32).pin_memory()
done_h = torch.empty((n, 1), dtype=torch.uint8).pin_memory()
a_dev = torch.empty(n, dtype=torch.float32, device="cuda")


def run_ranges(env, ranges, streams, zero_copy_actions=False):
    eng = env._engine
    cur = torch.cuda.current_stream()
    start = torch.cuda.Event()
    start.record(cur)
    for (lo, hi), s in zip(ranges, streams):
        with torch.cuda.stream(s):
            s.wait_event(start)
            if zero_copy_actions:
                eng.step_range(lo, hi, act, rew_h, done_h, obs_h)
            else:
                a_dev[lo:hi].copy_(act[lo:hi], non_blocking=True)
                eng.step_range(lo, hi, a_dev, rew_h, done_h, obs_h)
    for s in streams:
        cur.wait_stream(s)
    eng.flip()
    cur.synchronize()


for label, fr, zc in (("B 1 range", [1.0], False), ("C 2 ranges [1/8,7/8]", [0.125, 0.875], False),
                      ("C 4 ranges", [0.0625, 0.3125, 0.3125, 0.3125], False), ("C 8 ranges", [0.03125] + [0.96875 / 7] * 7, False)):
    try:
        env = make()
        edges = [0]
        acc = 0.0
        for f in fr:
            acc += f
            edges.append(min(n, int(round(acc * n)) // 16 * 16))
        edges[-1] = n
        ranges = [(edges[i], edges[i + 1]) for i in range(len(fr))]
        streams = [torch.cuda.Stream() for _ in ranges]
        dt = timeit(lambda: run_ranges(env, ranges, streams, zc))
        # correctness: same state history as mode A needs the same number of steps from the same start
        env2 = make()
        run_ranges(env2, ranges, streams, zc)
        envA = make()
        o, r, d = envA.step_host(act)[:3]
        ok = np.array_equal(obs_h.numpy(), o) and np.array_equal(rew_h.numpy(), r) and np.array_equal(done_h.numpy().astype(bool), d)
        print(f"{label:31s}: {dt*1e3:.3f} ms/step  {n/dt/1e9:.2f} G env-steps/s  equal to step_host: {ok}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"{label}: FAILED {type(ex).__name__}: {ex}", flush=True)
