"""step_host probe (development): where a C2 host step spends its 0.55 ms; chunk-count sweep."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import emei_b200 as E  # noqa: E402
from emei_b200.engine import HostStaging

n = 1 << 20
rng = np.random.default_rng(0)
st = (rng.uniform(-1, 1, size=(n, 4)) * np.array([4.0, 5.0, np.pi, 8.0])).astype(np.float32)
act = torch.as_tensor(rng.uniform(-1, 1, size=n).astype(np.float32)).pin_memory()
for chunks in (1, 2, 4, None, [0.25, 0.75], [0.125, 0.375, 0.5], [0.125, 0.875], [0.0625, 0.1875, 0.75], [0.1, 0.3, 0.6], [0.34, 0.66]):
    env = E.make("ContinuousCartPoleSwingUp-v0", freq_rate=4, num_envs=n, dtype=torch.float32)
    env.state = st
    env._staging = HostStaging(env, fractions=chunks) if isinstance(chunks, list) else HostStaging(env, chunks=chunks)
    for _ in range(5):
        env.step_host(act)
    t0 = time.perf_counter()
    reps = 40
    for _ in range(reps):
        env.step_host(act)
    dt = (time.perf_counter() - t0) / reps
    print(f"{str(chunks):24s} ranges={len(env._staging.ranges):2d} graphs={len(env._staging._graphs)}: {dt*1e3:.3f} ms/step  {n/dt/1e9:.2f} G env-steps/s")
