"""PCIe probe (development): pinned D2H / H2D copy bandwidth at the C2 step_host sizes, alone and concurrently."""
import time

import torch

dev = torch.device("cuda:0")
n = 1 << 20
obs_d = torch.empty((n, 4), dtype=torch.float32, device=dev)
rew_d = torch.empty((n, 1), dtype=torch.float32, device=dev)
done_d = torch.empty((n, 1), dtype=torch.uint8, device=dev)
obs_h, rew_h, done_h = (torch.empty_like(t, device="cpu").pin_memory() for t in (obs_d, rew_d, done_d))
act_h = torch.empty((n,), dtype=torch.float32).pin_memory()
act_d = torch.empty((n,), dtype=torch.float32, device=dev)
s2 = torch.cuda.Stream(dev)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def d2h():
    obs_h.copy_(obs_d, non_blocking=True)
    rew_h.copy_(rew_d, non_blocking=True)
    done_h.copy_(done_d, non_blocking=True)


def h2d():
    act_d.copy_(act_h, non_blocking=True)


def both():
    with torch.cuda.stream(s2):
        act_d.copy_(act_h, non_blocking=True)
    d2h()


b_d2h = sum(t.numel() * t.element_size() for t in (obs_h, rew_h, done_h))
t = timeit(d2h)
print(f"D2H {b_d2h/1e6:.1f} MB: {t*1e3:.3f} ms  {b_d2h/t/1e9:.1f} GB/s")
t = timeit(h2d)
print(f"H2D 4.2 MB: {t*1e3:.3f} ms  {act_h.numel()*4/t/1e9:.1f} GB/s")
t = timeit(both)
print(f"both directions: {t*1e3:.3f} ms per step -> {n/t/1e9:.2f} G env-steps/s bound")
big_h = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
big_d = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
t = timeit(lambda: big_h.copy_(big_d, non_blocking=True), 10)
print(f"D2H 268 MB: {(1<<28)/t/1e9:.1f} GB/s")
t = timeit(lambda: big_d.copy_(big_h, non_blocking=True), 10)
print(f"H2D 268 MB: {(1<<28)/t/1e9:.1f} GB/s")
