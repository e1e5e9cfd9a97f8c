"""Small-batch host step (development): kernel-side loads / stores of the pinned host arrays (zero copy) against the
copy-engine path, per batch size.  usage: python scripts/probe_zero_copy_small.py"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
if len(sys.argv) > 1:  # child: one mode
    sys.path.insert(0, ROOT)
    import torch

    import emei_b200 as E

    for env_id in ("BoundaryInvertedPendulumSwingUp-v0", "ContinuousCartPoleSwingUp-v0", "ChargedBallCentering-v0"):
        for n in (1024, 4096, 16384, 65536, 262144):
            env = E.make(env_id, num_envs=n, dtype=torch.float32, freq_rate=1)
            env.reset(seed=1)
            cont = len(env.action_space.shape) > 0
            rng = np.random.default_rng(0)
            act = torch.as_tensor(rng.uniform(-1, 1, size=n).astype(np.float32) if cont else rng.integers(0, 2, size=n).astype(np.uint8)).pin_memory()
            for _ in range(20):
                env.step_host(act)
            reps = 300
            t0 = time.perf_counter()
            for _ in range(reps):
                env.step_host(act)
            dt = (time.perf_counter() - t0) / reps
            print(f"{sys.argv[1]:10s} {env_id:40s} n={n:7d}  {dt*1e6:8.1f} us/step  {n/dt/1e6:9.1f} M env-steps/s  zero_copy={env._staging.zero_copy}", flush=True)
else:
    for mode, zc in (("dma", "0"), ("zero-copy", "1000000000")):
        subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, EMEI_ZERO_COPY_MAX=zc), check=False)
