#!/usr/bin/env python
"""bench.py -- headline benchmark of the emei_b200 hot path (contract: see the task statement).

Workload (BASELINE.json configs[1], "C2"): ContinuousCartPoleSwingUp batched step, 2^20 envs per
GPU, freq_rate=4 forward-Euler sub-steps, float32.  A "step" is one launch of the step kernel over
one 2^20-env batch.  Batches rotate over a ring whose footprint exceeds L2 (config.l2_policy), so
every timed launch streams its state from HBM.

  python bench.py [--gpus N --steps K --warmup W]            # our arm (torchrun for N>1)
  python bench.py --impl reference [--steps K --warmup W]    # CPU arm: the oracle port on host cores

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS = 1 << 20
FREQ_RATE = 4
DT = 0.02
ALG_BYTES_PER_ENV_STEP = 41  # state 16 + action 4 + next 16 + reward 4 + done 1
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = "ContinuousCartPoleSwingUp batched step, 2^20 envs/GPU, freq_rate=4, float32 (BASELINE configs[1])"


def synth_inputs(n, seed=1002):
    """[x, x', th, th'] = U(-1,1)*[4,5,pi,8] (+1% slice with |x| at the terminal
    threshold), action U(-1,1) float32 [n,1]."""
    rng = np.random.default_rng(seed)
    st = (rng.uniform(-1, 1, size=(n, 4)) * np.array([4.0, 5.0, np.pi, 8.0])).astype(np.float32)
    k = n // 100
    st[:k, 0] = np.sign(st[:k, 0]) * rng.uniform(4.99, 5.01, size=k).astype(np.float32)
    act = rng.uniform(-1, 1, size=(n, 1)).astype(np.float32)
    return st, act


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy restatement of the reference's step) on the host cores
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    st, act, reps = args
    from oracle import emei_oracle as O

    p = O.cartpole_params("continuous_swingup")
    t0 = time.perf_counter()
    for _ in range(reps):
        force = O.cartpole_force(act, True, p)
        nxt = O.cartpole_step_f64ref(st, force, DT, FREQ_RATE, p, libm=False)
        O.cartpole_reward("continuous_swingup", nxt)
        O.cartpole_terminal("continuous_swingup", nxt, p)
    return time.perf_counter() - t0


def cpu_step_rate(sample_envs, steps, cores):
    """env-steps/s of the oracle port: `cores` processes, each stepping sample_envs/cores envs."""
    import multiprocessing as mp

    st, act = synth_inputs(sample_envs)
    st = st.astype(np.float64)
    chunks = [(st[i::cores].copy(), act[i::cores].copy(), steps) for i in range(cores)]
    if cores == 1:
        t0 = time.perf_counter()
        _cpu_worker(chunks[0])
        wall = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, [(c[0][:1024], c[1][:1024], 1) for c in chunks])  # spin the workers up
            t0 = time.perf_counter()
            pool.map(_cpu_worker, chunks)
            wall = time.perf_counter() - t0
    return sample_envs * steps / wall, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    # calibrate, then size the per-step sample so the whole run stays within ~2 minutes
    rate1, _ = cpu_step_rate(1 << 16, 1, 1)
    budget_s = 100.0
    total_steps = args.steps + args.warmup
    sample = int(min(N_ENVS, max(1 << 12, rate1 * cores * 0.6 * budget_s / total_steps)))
    sample -= sample % cores
    if args.warmup:
        cpu_step_rate(sample, args.warmup, cores)
    rate, wall = cpu_step_rate(sample, args.steps, cores)
    desc = f"{sample} envs/step x {args.steps} steps of the numpy oracle port (oracle/emei_oracle.py), {cores} processes"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": rate,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_envs_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            for ln in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in ln.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    import emei_b200 as E
    from emei_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; emei_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, args.warmup
    ring = args.ring
    st, act = synth_inputs(N_ENVS, seed=1002 + rank)
    # ring of independent 2^20-env batches: 41 MB of step traffic each, ring*41 MB >> 126 MB of L2
    envs, acts = [], []
    for j in range(ring):
        env = E.make("ContinuousCartPoleSwingUp-v0", freq_rate=FREQ_RATE, real_time_scale=DT, num_envs=N_ENVS,
                     dtype=torch.float32, device=dev, env_offset=(rank * ring + j) * N_ENVS)
        env.state = np.roll(st, j * 4099, axis=0)
        env._stats = envs[0].stats if envs else env.stats  # one shared statistics buffer
        envs.append(env)
        acts.append(torch.as_tensor(np.roll(act, j * 4099, axis=0)).to(dev).reshape(N_ENVS))
    envs[0].reset_stats()

    def one_step(i):
        envs[i % ring].step(acts[i % ring])

    # ---------------- device-resident value: W warm-up steps, then EXACTLY K steps in one CUDA graph
    for i in range(max(W, 3)):
        one_step(i)
    torch.cuda.synchronize()
    launches0 = _lib.launch_count
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(K):
                one_step(i)
    launches_per_replay = _lib.launch_count - launches0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    graph.replay()  # warm the instantiated graph (also puts the GPU under load for the clock samples)
    torch.cuda.synchronize()
    stats_buf = envs[0].stats
    envs[0].reset_stats()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 1  # EXACTLY K steps are timed
    ev0.record()
    for _ in range(reps):
        graph.replay()
    ev1.record()
    if world > 1:
        dist.all_reduce(stats_buf)  # end-of-rollout statistics: 2 doubles over NCCL/NVLink
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1) / reps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = world * N_ENVS * K / (ms_total * 1e-3)

    # ---------------- e2e: host actions in, host obs/reward/done out, every step (public step_host API)
    e2e_steps = max(3, min(K, args.e2e_steps))
    act_host = [torch.as_tensor(np.roll(act, j * 4099, axis=0).reshape(N_ENVS)).pin_memory() for j in range(min(ring, 4))]
    env0 = envs[0]
    for i in range(3):
        env0.step_host(act_host[i % len(act_host)])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        env0.step_host(act_host[i % len(act_host)])
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N_ENVS * e2e_steps / (float(te.item()) * 1e-3)
    clocks = sampler.finish() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    achieved = ALG_BYTES_PER_ENV_STEP * N_ENVS / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": K,
        "warmup": max(W, 3),
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "envs_per_gpu": N_ENVS,
            "freq_rate": FREQ_RATE,
            "real_time_scale": DT,
            "l2_policy": f"inputs larger than L2: ring of {ring} independent 2^20-env batches ({ring * 41} MB of step traffic) rotated every launch",
            "launch": f"K={K} step launches captured in one CUDA graph, replayed once; CUDA events on the launching stream",
            "parallelism": f"env batch sharded, {world} rank(s), no data-path collective",
        },
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": None,
            "peak_source": peak_src,
            "algorithmic_bytes_per_env_step": ALG_BYTES_PER_ENV_STEP,
            "kernel": "emei::cartpole_step_kernel<float,false>",
            "note": "duration = timed region / launches (includes inter-launch gaps)",
        },
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": env0._staging.h2d_bytes,
            "d2h_bytes_per_step": env0._staging.d2h_bytes,
            "steps": e2e_steps,
            "api": "env.step_host(action_host) -> numpy obs/reward/terminated (pinned staging, sync per step)",
        },
        "gpu_launches": launches_per_replay * reps,
        "clocks": clocks,
    }
    # ---------------- CPU baseline on this box's host cores (bounded sample; oracle = checker only)
    if world == 1 and not args.no_cpu:
        rate1, wall1 = cpu_step_rate(1 << 20, 16, 1)
        line["cpu_baseline"] = {
            "value": rate1,
            "unit": UNIT,
            "cores": 1,
            "kind": "port",
            "sample": f"2^20 envs x 16 steps of the numpy oracle port (vectorised restatement of base_control.py:61-83), {wall1:.1f} s",
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
