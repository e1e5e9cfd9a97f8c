import torch, sys, time
sys.path.insert(0, '.')
import emei_b200 as E
n = 1 << 26
env = E.make("ChargedBallCentering-v0", num_envs=n, dtype=torch.float32, device="cuda:0")
env.reset(seed=1004)
g = torch.Generator(device="cuda:0"); g.manual_seed(1004)
acts = [torch.randint(0, 2, (n,), device="cuda:0", dtype=torch.uint8, generator=g) for _ in range(4)]
evs = [torch.cuda.Event(enable_timing=True) for _ in range(81)]
torch.cuda.synchronize()
evs[0].record()
for t in range(80):
    env.step(acts[t % 4])
    evs[t + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(80)]
print(" ".join(f"{x:.2f}" for x in ts))
print("on_circle frac", float(env.state["on_circle"].float().mean()))
