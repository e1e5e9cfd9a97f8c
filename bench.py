#!/usr/bin/env python
"""bench.py -- benchmark of the emei_b200 hot path (contract: see the task statement / DESIGN.md 6).

Default workload = BASELINE.json configs[1] ("c2"): ContinuousCartPoleSwingUp batched step, 2^20 envs
per GPU, freq_rate=4 forward-Euler sub-steps, float32.  A "step" is one pass of the hot path over
one batch.  The other BASELINE configs are selectable with --workload (they are parity-test sizes
and secondary bench lines, committed under profiles/):

  c1            BoundaryInvertedPendulumSwingUp step, 4096 envs, freq_rate=1   (launch-latency bound)
  c2 (default)  ContinuousCartPoleSwingUp step, 2^20 envs/GPU, freq_rate=4
  c3_hopper     Hopper get_batch_reward+get_batch_terminal, 2^24 transitions/GPU (terminate_when_unhealthy=False)
  c3_halfcheetah  HalfCheetah same, 2^24 transitions/GPU
  c4            ChargedBallCentering step, 2^26 envs TOTAL sharded over the ranks (strong scaling)
  c5            scoring 2^26 transitions/GPU (half Hopper, half HalfCheetah) + freeze/unfreeze of a
                2^26-env cart-pole state buffer per sweep

  python bench.py [--gpus N --steps K --warmup W] [--workload W]          # our arm (torchrun for N>1)
  python bench.py --impl reference [--steps K --warmup W] [--workload W]  # CPU arm: the oracle port on host cores

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

DT = 0.02


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ==================================================================================================
# synthetic inputs: generated on the host, seeded, identical bits for CPU and GPU arms
# ==================================================================================================
def synth_cartpole(n, seed=1002):
    """C2: [x, x', th, th'] = U(-1,1)*[4,5,pi,8] (+1% slice with |x| at the terminal threshold), action U(-1,1)."""
    rng = np.random.default_rng(seed)
    st = (rng.uniform(-1, 1, size=(n, 4)) * np.array([4.0, 5.0, np.pi, 8.0])).astype(np.float32)
    k = n // 100
    st[:k, 0] = np.sign(st[:k, 0]) * rng.uniform(4.99, 5.01, size=k).astype(np.float32)
    act = rng.uniform(-1, 1, size=(n, 1)).astype(np.float32)
    return st, act


def synth_ip(n, seed=1001):
    """C1: [x, th, v, w] = U(-1,1)*[1.9, pi, 5, 8], ctrl U(-3,3)."""
    rng = np.random.default_rng(seed)
    st = (rng.uniform(-1, 1, size=(n, 4)) * np.array([1.9, np.pi, 5.0, 8.0])).astype(np.float32)
    act = rng.uniform(-3, 3, size=(n, 1)).astype(np.float32)
    return st, act


def synth_i2p(n, seed=1006):
    """I2P: [x, th0, th1, v, w0, w1] = U(-1,1)*[2.9, pi, pi, 4, 8, 10], ctrl U(-1,1) (tests/test_gpu_parity.py i2p_inputs)."""
    rng = np.random.default_rng(seed)
    st = (rng.uniform(-1, 1, size=(n, 6)) * np.array([2.9, np.pi, np.pi, 4.0, 8.0, 10.0])).astype(np.float32)
    act = rng.uniform(-1, 1, size=(n, 1)).astype(np.float32)
    return st, act


def synth_scoring(n, family, seed=1003):
    """C3: Hopper obs = [0,1.25,0..] + N(0,1)*[1,.4,.15,1..]; HalfCheetah obs N(0,1)[n,18]; pre_obs = obs with
    column 0 shifted by N(0, 0.01); action U(-1,1); 0.1% rows poisoned with NaN/Inf."""
    rng = np.random.default_rng(seed + (0 if family == "hopper" else 1))
    d, da = (12, 3) if family == "hopper" else (18, 6)
    obs = rng.standard_normal((n, d), dtype=np.float32)
    if family == "hopper":
        obs *= np.array([1, 0.4, 0.15] + [1] * 9, dtype=np.float32)
        obs[:, 1] += 1.25
    pre = obs.copy()
    pre[:, 0] -= 0.01 * rng.standard_normal(n, dtype=np.float32)
    act = rng.uniform(-1, 1, size=(n, da)).astype(np.float32)
    bad = rng.integers(0, n, size=max(1, n // 1000))
    obs[bad, rng.integers(0, d, size=bad.shape[0])] = rng.choice(np.array([np.nan, np.inf, -np.inf], dtype=np.float32), size=bad.shape[0])
    return obs, pre, act


# ==================================================================================================
# CPU arm: the oracle port (numpy restatement of the reference) on the host cores.  bench.py's
# cpu_baseline / --impl reference are the ONLY product-side users of oracle/ (as the thing timed
# beside the GPU path, never inside it).
# ==================================================================================================
def cpu_port_name():
    """Which restatement the CPU legs time: the plain-C one (oracle/emei_oracle_c.c, built by build()) where it covers
    the workload and its library is present, else the numpy one."""
    try:
        from oracle import c_oracle as C

        return "C" if C.available() else "numpy"
    except Exception:
        return "numpy"


def cpu_port_desc(kind):
    if cpu_port_name() == "C" and kind in ("c2", "c1", "i2p", "c3_hopper", "c3_halfcheetah", "c4"):
        return f"plain-C oracle port (oracle/emei_oracle_c.c, kind={kind})"
    return f"numpy oracle port (oracle/emei_oracle.py, kind={kind})"


def _cpu_worker(args):
    kind, arrays, reps = args
    from oracle import emei_oracle as O

    C = None
    if cpu_port_name() == "C":
        from oracle import c_oracle as C

    t0 = time.perf_counter()
    for _ in range(reps):
        if kind == "c2":
            st, act = arrays
            p = O.cartpole_params("continuous_swingup")
            if C is not None:
                nxt = C.cartpole_step_f64ref(st, O.cartpole_force(act, True, p), DT, 4, p)
                C.cartpole_reward_terminal("continuous_swingup", nxt, p)
                continue
            nxt = O.cartpole_step_f64ref(st, O.cartpole_force(act, True, p), DT, 4, p, libm=False)
            O.cartpole_reward("continuous_swingup", nxt)
            O.cartpole_terminal("continuous_swingup", nxt, p)
        elif kind == "c1":
            st, act = arrays
            p = O.InvertedPendulumParams()
            if C is not None:
                nxt, obs = C.ip_step(st, np.clip(act[:, 0].astype(np.float64), p.ctrl_low, p.ctrl_high), DT, 1, True, p)
            else:
                nxt, obs = O.ip_step(st, act[:, 0].astype(np.float64), DT, 1, True, p)
            O.ip_reward("ip_boundary_swingup", obs)
            O.ip_terminal("ip_boundary_swingup", obs, p)
        elif kind == "i2p":
            st, act = arrays
            p = O.I2PParams()
            if C is not None:
                nxt, obs = C.i2p_step(st, act[:, 0].astype(np.float64), DT, 1, True, p)
            else:
                nxt, obs = O.i2p_step(st, act[:, 0].astype(np.float64), DT, 1, True, p)
            O.i2p_reward("i2p_boundary_swingup", obs)
            O.i2p_terminal("i2p_boundary_swingup", obs)
        elif kind == "c3_hopper":
            obs, pre, act = arrays
            p = O.HopperParams(terminate_when_unhealthy=False)
            if C is not None:
                C.hopper_reward_terminal(obs, pre, act, p)
                continue
            O.hopper_reward(obs, pre, act, p)
            O.hopper_terminal(obs, p)
        elif kind == "c3_halfcheetah":
            obs, pre, act = arrays
            p = O.HalfCheetahParams()
            if C is not None:
                C.halfcheetah_reward_terminal(obs, pre, act, p)
                continue
            O.halfcheetah_reward(obs, pre, act, p)
            O.halfcheetah_terminal(obs)
        elif kind == "c4":
            on, ci, fr, act = arrays
            p = O.ChargedBallParams()
            if C is not None:
                on, ci, fr = C.charged_ball_step(on, ci, fr, O.charged_ball_force(act, False, p), 1, p)
                C.charged_ball_reward(fr, p)
                continue
            on, ci, fr = O.charged_ball_step(on, ci, fr, O.charged_ball_force(act, False, p), 1, p)
            O.charged_ball_reward(fr, p)
        else:
            raise ValueError(kind)
    return time.perf_counter() - t0


def _cpu_inputs(kind, n):
    if kind == "c2":
        st, act = synth_cartpole(n)
        return [st.astype(np.float64), act]
    if kind == "c1":
        st, act = synth_ip(n)
        return [st.astype(np.float64), act]
    if kind == "i2p":
        st, act = synth_i2p(n)
        return [st.astype(np.float64), act]
    if kind in ("c3_hopper", "c3_halfcheetah"):
        obs, pre, act = synth_scoring(n, kind[3:])
        return [obs.astype(np.float64), pre.astype(np.float64), act.astype(np.float64)]
    if kind == "c4":
        from oracle import emei_oracle as O

        rng = np.random.default_rng(1004)
        on, ci, fr = O.charged_ball_init_state(n, rng, O.ChargedBallParams())
        return [on, ci, fr, rng.integers(0, 2, size=n)]
    raise ValueError(kind)


_CPU_CHUNKS = None


def _cpu_worker_indexed(i):
    return _cpu_worker(_CPU_CHUNKS[i])


def _cpu_worker_warm(i):
    kind, arrays, _ = _CPU_CHUNKS[i]
    return _cpu_worker((kind, [a[:256] for a in arrays], 1))


def cpu_rate(kind, sample_units, reps, cores):
    """units/s of the oracle port: `cores` processes, each owning sample_units/cores units."""
    import multiprocessing as mp

    global _CPU_CHUNKS
    arrays = _cpu_inputs(kind, sample_units)
    chunks = [(kind, [a[i::cores].copy() for a in arrays], reps) for i in range(cores)]
    with np.errstate(all="ignore"):
        if cores == 1:
            t0 = time.perf_counter()
            _cpu_worker(chunks[0])
            wall = time.perf_counter() - t0
        else:
            # the workers are forked AFTER the inputs exist and inherit them (copy-on-write): the timed map ships only
            # an index, not gigabytes of pickled arrays, so the CPU arm is timed on its arithmetic
            _CPU_CHUNKS = chunks
            with mp.get_context("fork").Pool(cores) as pool:
                pool.map(_cpu_worker_warm, range(cores))  # spin the workers up, touch the inherited pages
                t0 = time.perf_counter()
                pool.map(_cpu_worker_indexed, range(cores))
                wall = time.perf_counter() - t0
            _CPU_CHUNKS = None
    return sample_units * reps / wall, wall


# ==================================================================================================
# workloads
# ==================================================================================================
class Workload:
    key = name = metric = unit = kernel = ""
    dtype = "f32"
    scaling = "weak"
    use_graph = True
    bound = "hbm"  # which roofline bounds the dominant kernel: "hbm" | "issue" (warp-instruction issue)
    alg_bytes = 0  # algorithmic bytes per unit (SURVEY.md 8d / DESIGN.md 5)
    cpu_kind = None  # which oracle routine is the CPU baseline
    cpu_sample = 1 << 20

    def __init__(self, args, rank, world, dev):
        self.args, self.rank, self.world, self.dev = args, rank, world, dev

    # units processed by THIS rank in one step
    units = 0

    def setup(self):
        raise NotImplementedError

    def step(self, i):
        raise NotImplementedError

    def setup_e2e(self):
        raise NotImplementedError

    def step_e2e(self, i):
        raise NotImplementedError

    def config(self):
        return {}

    def dominant_kernel_ms(self, ms_per_step):
        """duration of the dominant kernel per launch; default: the whole step is that one kernel."""
        return ms_per_step


class CartPoleStep(Workload):
    """C2 (and, with the IP env, C1): ring of independent batches larger than L2, rotated every launch."""

    key, metric, unit = "c2", "env_steps_per_sec", "env-steps/s"
    name = "ContinuousCartPoleSwingUp batched step, 2^20 envs/GPU, freq_rate=4, float32 (BASELINE configs[1])"
    kernel = "emei::cartpole_step_f32_tma_kernel<IP=0, AK=f32, FR=4, HAS_OBS=0>"
    env_id, n_envs, freq_rate, ring = "ContinuousCartPoleSwingUp-v0", 1 << 20, 4, 8
    alg_bytes = 41  # state 16 + action 4 + next 16 + reward 4 + done 1
    cpu_kind = "c2"
    inst_per_unit = 179.1  # thread-level SASS instructions per env-step = 32 x smsp__inst_executed / envs (packed f32x2: one FFMA2 serves two envs); ncu, profiles/r01_launches_bench_c2.csv

    def synth(self, seed):
        return synth_cartpole(self.n_envs, seed)

    def setup(self):
        import torch

        import emei_b200 as E

        self.units = self.n_envs
        ring = self.args.ring or self.ring
        st, act = self.synth(1002 + self.rank)
        self.envs, self.acts = [], []
        for j in range(ring):
            env = E.make(self.env_id, freq_rate=self.freq_rate, real_time_scale=DT, num_envs=self.n_envs,
                         dtype=torch.float32, device=self.dev, env_offset=(self.rank * ring + j) * self.n_envs)
            env.state = np.roll(st, j * 4099, axis=0)
            env._stats = self.envs[0].stats if self.envs else env.stats  # one shared statistics buffer
            self.envs.append(env)
            self.acts.append(torch.as_tensor(np.roll(act, j * 4099, axis=0)).to(self.dev).reshape(self.n_envs))
        self.envs[0].reset_stats()
        self.stats = self.envs[0].stats
        self._act_host = act

    def step(self, i):
        j = i % len(self.envs)
        self.envs[j].step(self.acts[j])

    def setup_e2e(self):
        import torch

        a = self._act_host
        self.act_host = [torch.as_tensor(np.roll(a, j * 4099, axis=0).reshape(self.n_envs)).pin_memory() for j in range(4)]
        self.env0 = self.envs[0]
        self.env0.step_host(self.act_host[0])
        self.h2d, self.d2h = self.env0._staging.h2d_bytes, self.env0._staging.d2h_bytes
        self.e2e_api = "env.step_host(action_host) -> numpy obs/reward/terminated (pinned staging, chunked copy/compute overlap)"

    e2e_warmup = 24  # 4 pinned action buffers x 2 ping-pong sides, each seen eagerly, captured, replayed

    def step_e2e(self, i):
        self.env0.step_host(self.act_host[i % len(self.act_host)])

    def config(self):
        ring = len(self.envs)
        mb = ring * self.n_envs * self.alg_bytes / 1e6
        return {
            "envs_per_gpu": self.n_envs, "freq_rate": self.freq_rate, "real_time_scale": DT,
            "l2_policy": f"inputs larger than L2: ring of {ring} independent {self.n_envs}-env batches ({mb:.0f} MB of step traffic) rotated every launch",
        }


class IPStep(CartPoleStep):
    key = "c1"
    name = "BoundaryInvertedPendulumSwingUp batched step, 4096 envs, freq_rate=1, float32 (BASELINE configs[0])"
    kernel = "emei::cartpole_step_f32_small_kernel<IP=1, AK=f32, FR=1, HAS_OBS=1>"
    env_id, n_envs, freq_rate, ring = "BoundaryInvertedPendulumSwingUp-v0", 4096, 1, 1024
    alg_bytes = 57  # state 16 + action 4 + next state 16 + wrapped obs 16 + reward 4 + done 1
    cpu_kind, cpu_sample = "c1", 1 << 20
    inst_per_unit = 202.5  # one env per thread, 128-thread CTAs: profiles/r01_launches_bench_c1.csv

    def synth(self, seed):
        return synth_ip(self.n_envs, seed)


class CartPoleStepLarge(CartPoleStep):
    """The C2 step at 2^24 envs per GPU: launch ramp and drain amortised, the kernel's steady state (recycled TMA ring)."""

    key = "c2_large"
    name = "ContinuousCartPoleSwingUp batched step, 2^24 envs/GPU, freq_rate=4, float32 (C2's kernel at 16x the batch)"
    n_envs, ring = 1 << 24, 2
    use_graph = False


class I2PStep(CartPoleStep):
    """the analytic inverted double pendulum step (dynamics + observation + reward/terminal in one
    launch), same protocol as C2."""

    key = "i2p"
    name = "BoundaryInvertedDoublePendulumSwingUp batched step, 2^20 envs/GPU, freq_rate=1, float32 (SURVEY 8f rank 3)"
    kernel = "emei::i2p_step_kernel<float>"
    env_id, n_envs, freq_rate, ring = "BoundaryInvertedDoublePendulumSwingUp-v0", 1 << 20, 1, 8
    alg_bytes = 81  # state 24 + action 4 + next state 24 + observation 24 + reward 4 + done 1
    cpu_kind, cpu_sample = "i2p", 1 << 18
    inst_per_unit = 342.7  # 32 x smsp__inst_executed / envs (ncu, profiles/r01_launches_i2p.csv)

    def synth(self, seed):
        return synth_i2p(self.n_envs, seed)

    def setup_e2e(self):
        self.h2d = self.d2h = 0
        self.e2e_api = None  # step_host covers the cart-pole / IP / charged-ball engines


class Scoring(Workload):
    """C3: fused get_batch_reward + get_batch_terminal over n transitions resident in HBM (>> L2)."""

    metric, unit = "transitions_per_sec", "transitions/s"
    use_graph = False
    family, n = "hopper", 1 << 24
    env_kwargs = {}

    def setup(self):
        import torch

        import emei_b200 as E

        self.units = self.n
        self.env = E.make(self.env_id, dtype=torch.float32, device=self.dev, **self.env_kwargs)
        obs, pre, act = synth_scoring(self.n, self.family, 1003 + 17 * self.rank)
        self._host = (obs, pre, act)
        self.obs, self.pre, self.act = (torch.as_tensor(a).to(self.dev) for a in (obs, pre, act))
        self.stats = self.env.stats
        self._ev = []

    def step(self, i):
        self.out = self.env.get_batch_reward_terminal(self.obs, self.pre, self.act)

    def setup_e2e(self):
        import torch

        self.h_in = [torch.as_tensor(a).pin_memory() for a in self._host]
        self.h_r = torch.empty((self.n, 1), dtype=torch.float32).pin_memory()
        self.h_d = torch.empty((self.n, 1), dtype=torch.bool).pin_memory()
        self.h2d = sum(t.numel() * t.element_size() for t in self.h_in)
        self.d2h = self.h_r.numel() * 4 + self.h_d.numel()
        self.e2e_api = "env.get_batch_reward_terminal(obs, pre_obs, action) with pinned HOST tensors -> reward/terminal copied back to pinned host"

    def step_e2e(self, i):
        import torch

        r, d = self.env.get_batch_reward_terminal(*self.h_in)
        self.h_r.copy_(r, non_blocking=True)
        self.h_d.copy_(d, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()

    def config(self):
        return {"transitions_per_gpu": self.n, "l2_policy": f"inputs larger than L2 ({self.n * self.alg_bytes / 1e6:.0f} MB per sweep)",
                "note": "step = emei_sumsq (batch-wide control cost, 1 launch) + fused emei_reward_terminal (1 launch)"}


class HopperScoring(Scoring):
    key, family, env_id = "c3_hopper", "hopper", "HopperRunning-v0"
    name = "Hopper get_batch_reward+get_batch_terminal, 2^24 transitions/GPU, terminate_when_unhealthy=False, float32 (BASELINE configs[2])"
    kernel = "emei::reward_terminal_kernel<float, HOPPER> (+ emei::sumsq_kernel<float>)"
    env_kwargs = {"terminate_when_unhealthy": False}
    alg_bytes = 69  # obs 48 + pre_obs[:,0] 4 + action 12 + reward 4 + done 1
    cpu_kind, cpu_sample = "c3_hopper", 1 << 22


class HalfCheetahScoring(Scoring):
    key, family, env_id = "c3_halfcheetah", "halfcheetah", "HalfCheetahRunning-v0"
    name = "HalfCheetah get_batch_reward+get_batch_terminal, 2^24 transitions/GPU, float32 (BASELINE configs[2])"
    kernel = "emei::reward_terminal_kernel<float, HALFCHEETAH> (+ emei::sumsq_kernel<float>)"
    alg_bytes = 105  # obs 72 + pre_obs[:,0] 4 + action 24 + reward 4 + done 1
    cpu_kind, cpu_sample = "c3_halfcheetah", 1 << 22


class ChargedBall(Workload):
    """C4: 2^26 envs in total, sharded contiguously over the ranks (strong scaling); state updated in place."""

    key, metric, unit = "c4", "env_steps_per_sec", "env-steps/s"
    name = "ChargedBallCentering batched step, 2^26 envs total sharded over the ranks, freq_rate=1, float32 (BASELINE configs[3])"
    kernel = "emei::charged_ball_step_f32_kernel<AK=u8>"
    scaling, use_graph = "strong", False
    total = 1 << 26
    alg_bytes = 56  # state in 25 + action 1 (uint8) + state out 25 + reward 4 + done 1
    cpu_kind, cpu_sample = "c4", 1 << 20

    def setup(self):
        import torch

        import emei_b200 as E
        from emei_b200.dist import shard_range

        if self.args.total_log2:
            self.total = 1 << self.args.total_log2
        b, e = shard_range(self.total, self.rank, self.world)
        self.units = e - b
        self.env = E.make("ChargedBallCentering-v0", num_envs=self.units, dtype=torch.float32, device=self.dev, env_offset=b)
        self.env.reset(seed=1004)
        g = torch.Generator(device=self.dev)
        g.manual_seed(1004 + self.rank)
        self.acts = [torch.randint(0, 2, (self.units,), device=self.dev, dtype=torch.uint8, generator=g) for _ in range(4)]
        self.env.reset_stats()
        self.stats = self.env.stats

    def step(self, i):
        self.env.step(self.acts[i % 4])

    def setup_e2e(self):
        import torch

        self.act_host = [a.cpu().pin_memory() for a in self.acts]
        self.env.step_host(self.act_host[0])
        self.h2d, self.d2h = self.env._staging.h2d_bytes, self.env._staging.d2h_bytes
        self.e2e_api = "env.step_host(action_host) -> numpy obs/reward/terminated (pinned staging)"

    e2e_warmup = 8  # 4 pinned action buffers, each seen eagerly once, then captured

    def step_e2e(self, i):
        self.env.step_host(self.act_host[i % 4])

    def config(self):
        return {"envs_total": self.total, "envs_per_gpu": self.units, "freq_rate": 1,
                "l2_policy": f"inputs larger than L2 ({self.units * self.alg_bytes / 1e6:.0f} MB per step per GPU)"}


class ScoringSweep(Workload):
    """C5: one imagined-rollout scoring sweep per step = freeze() of a 2^26-env cart-pole state buffer,
    Hopper + HalfCheetah reward/terminal over 2^25 transitions each, unfreeze()."""

    key, metric, unit = "c5", "transitions_per_sec", "transitions/s"
    name = "MBRL scoring sweep: 2^26 transitions/GPU (half Hopper, half HalfCheetah) reward+terminal + freeze/unfreeze of a 2^26-env cart-pole buffer, float32 (BASELINE configs[4])"
    kernel = "emei::reward_terminal_kernel<float, HALFCHEETAH> (dominant), HOPPER, sumsq, snapshot_copy"
    use_graph = False
    n_half, n_env = 1 << 25, 1 << 26
    # per transition: (69 + 105)/2 scoring + 2 x (16 read + 16 write) snapshot bytes per env over n_env == 2*n_half transitions
    alg_bytes = (69 + 105) / 2 + 64
    cpu_kind, cpu_sample = "c3_hopper", 1 << 22

    def setup(self):
        import torch

        import emei_b200 as E

        self.units = 2 * self.n_half
        self.hop = E.make("HopperRunning-v0", terminate_when_unhealthy=False, dtype=torch.float32, device=self.dev)
        self.chee = E.make("HalfCheetahRunning-v0", dtype=torch.float32, device=self.dev)
        self.cp = E.make("CartPoleSwingUp-v0", num_envs=self.n_env, dtype=torch.float32, device=self.dev)
        self.cp.reset(seed=1005)
        self.data = []
        for fam in ("hopper", "halfcheetah"):
            obs, pre, act = synth_scoring(self.n_half, fam, 1005 + 31 * self.rank)
            self.data.append(tuple(torch.as_tensor(a).to(self.dev) for a in (obs, pre, act)))
        self.stats = self.hop.stats

    def step(self, i):
        self.cp.freeze()
        self.o1 = self.hop.get_batch_reward_terminal(*self.data[0])
        self.o2 = self.chee.get_batch_reward_terminal(*self.data[1])
        self.cp.unfreeze()

    def setup_e2e(self):
        self.h2d = self.d2h = 0
        self.e2e_api = None

    def config(self):
        return {"transitions_per_gpu": self.units, "snapshot_envs_per_gpu": self.n_env,
                "l2_policy": "inputs larger than L2 (10 GB per sweep)"}


class CartPoleRollout(Workload):
    """the collection loop (zoo/util.py:33-93) as ONE launch per `horizon` env-steps: state,
    TimeLimit counter and episode return in registers, in-kernel auto-reset and uniform random policy.  No
    per-step HBM traffic (24 B per env per launch), so the bound is warp-instruction issue."""

    key, metric, unit = "rollout", "env_steps_per_sec", "env-steps/s"
    name = "ContinuousCartPoleSwingUp fused rollout, 2^20 envs/GPU x horizon steps per launch, freq_rate=4, in-kernel random policy + TimeLimit(1000) + auto-reset, float32 (SURVEY 8f rank 1)"
    kernel = "emei::rollout_f32_kernel<CartPoleDyn<IP=0, AK=f32, FR=4>, RECORD=0>"
    env_id, n_envs, freq_rate = "ContinuousCartPoleSwingUp-v0", 1 << 20, 4
    use_graph, bound = False, "issue"
    record = False
    alg_bytes = 0.0  # set in setup(): per env-step
    inst_per_unit = 247.1  # thread-level SASS instructions per env-step incl. divergent in-kernel resets (32 x smsp__inst_executed / env-steps; ncu, profiles/r01_launches_rollout*.csv)
    cpu_kind = "c2"
    e2e_max_steps = 5

    def setup(self):
        import torch

        import emei_b200 as E

        self.T = self.args.horizon
        self.units = self.n_envs * self.T
        self.env = E.make(self.env_id, freq_rate=self.freq_rate, real_time_scale=DT, num_envs=self.n_envs,
                          dtype=torch.float32, device=self.dev, env_offset=self.rank * self.n_envs)
        self.env.reset(seed=1006)
        self.stats = self.env.stats
        # state 16 + 3 counters 12, read and written once per launch; records (if any) per env-step
        self.alg_bytes = 2 * 28 / self.T + (42 if self.record else 0)

    def step(self, i):
        self.out = self.env.rollout(self.T, record=self.record)

    def setup_e2e(self):
        import torch

        rng = np.random.default_rng(1006 + self.rank)
        self.act_host = [torch.as_tensor(rng.uniform(-1, 1, size=(self.T, self.n_envs)).astype(np.float32)).pin_memory() for _ in range(2)]
        self.h2d = self.act_host[0].numel() * 4
        self.d2h = 48
        self.e2e_api = "env.rollout(horizon, actions=pinned HOST [T,n] float32) -> env.rollout_info(stats) read on the host"

    def step_e2e(self, i):
        out = self.env.rollout(self.T, actions=self.act_host[i % 2])
        self.info = self.env.rollout_info(out["stats"])

    def config(self):
        return {"envs_per_gpu": self.n_envs, "horizon": self.T, "freq_rate": self.freq_rate, "max_episode_steps": 1000,
                "records": self.record,
                "l2_policy": "no reuse to defeat: every env's state is read once and written once per launch"
                             + (f"; records are {self.units * 42 / 1e6:.0f} MB of fresh writes per launch" if self.record else "")}


class CartPoleRolloutRecord(CartPoleRollout):
    """Same launch, additionally writing every transition in the reference's dataset layout (zoo/util.py:62-67):
    observations/next_observations [T,n,4], actions/rewards [T,n], dones/timeouts u8[T,n] = 42 B per env-step."""

    key = "rollout_rec"
    name = CartPoleRollout.name.replace("fused rollout", "fused rollout + transition records (dataset layout)")
    kernel = "emei::rollout_f32_kernel<CartPoleDyn<IP=0, AK=f32, FR=4>, RECORD=1>"
    record = True
    inst_per_unit = 289.9  # (ncu, profiles/r01_launches_rollout_rec.csv) t_issue = 0.26 ms > t_hbm = 0.22 ms (42 B/env-step) at 2^25 env-steps per launch: still issue-bound

    def setup(self):
        if not self.args.horizon_set:
            self.args.horizon = 32
        super().setup()

    def setup_e2e(self):
        self.h2d = self.d2h = 0
        self.e2e_api = None


class ChargedBallRollout(Workload):
    """C4 as BASELINE words it ("rollouts, 64M envs x 200 steps"): ONE launch advances every env of the shard by
    `horizon` steps with the state in registers (emei_charged_ball_rollout_f32), in-kernel Bernoulli(1/2) policy,
    TimeLimit(500) + auto-reset.  HBM is touched once per env per launch, so the bound is warp-instruction issue."""

    key, metric, unit = "c4_rollout", "env_steps_per_sec", "env-steps/s"
    name = "ChargedBallCentering fused rollouts, 2^26 envs total sharded over the ranks x 200 steps per launch, freq_rate=1, float32 (BASELINE configs[3])"
    kernel = "emei::rollout_f32_kernel<ChargedBallDyn<u8>, RECORD=0>"
    scaling, use_graph, bound = "strong", False, "issue"
    total = 1 << 26
    inst_per_unit = 166.0  # warp-level SASS instructions per env-step, all three divergent paths issued (ncu smsp__inst_executed, profiles/r01_launches_c4_rollout.csv)
    cpu_kind, cpu_sample = "c4", 1 << 20
    e2e_max_steps = 5

    def setup(self):
        import torch

        import emei_b200 as E
        from emei_b200.dist import shard_range

        if self.args.total_log2:
            self.total = 1 << self.args.total_log2
        self.T = self.args.horizon if self.args.horizon_set else 200
        b, e = shard_range(self.total, self.rank, self.world)
        self.n = e - b
        self.units = self.n * self.T
        self.env = E.make("ChargedBallCentering-v0", num_envs=self.n, dtype=torch.float32, device=self.dev, env_offset=b)
        self.env.reset(seed=1004)
        self.stats = self.env.stats
        self.alg_bytes = 2 * (25 + 12) / self.T  # state 25 + 3 episode counters 12, read + written once per launch

    def step(self, i):
        self.out = self.env.rollout(self.T)

    def setup_e2e(self):
        import torch

        # teacher-forced policy from the host: uint8 [T, n] actions in pinned memory, statistics read back
        self.Te = min(self.T, 25)
        g = torch.Generator()
        g.manual_seed(1004 + self.rank)
        self.act_host = [torch.randint(0, 2, (self.Te, self.n), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
        self.h2d, self.d2h = self.Te * self.n, 48
        self.e2e_units = self.n * self.Te
        self.e2e_api = f"env.rollout({self.Te}, actions=pinned HOST uint8[{self.Te}, n]) -> env.rollout_info(stats) read on the host"

    def step_e2e(self, i):
        out = self.env.rollout(self.Te, actions=self.act_host[i % 2])
        self.info = self.env.rollout_info(out["stats"])

    def config(self):
        return {"envs_total": self.total, "envs_per_gpu": self.n, "horizon": self.T, "freq_rate": 1, "max_episode_steps": 500,
                "l2_policy": f"no reuse to defeat: {self.n * 37 / 1e6:.0f} MB of state read once and written once per launch"}


WORKLOADS = {w.key: w for w in (IPStep, CartPoleStep, CartPoleStepLarge, I2PStep, HopperScoring, HalfCheetahScoring, ChargedBall, ScoringSweep,
                                CartPoleRollout, CartPoleRolloutRecord, ChargedBallRollout)}


# ==================================================================================================
# reference arm
# ==================================================================================================
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    cores = host_cores()
    kind = W.cpu_kind
    rate1, _ = cpu_rate(kind, 1 << 14, 1, 1)  # calibrate, then size the per-step sample for ~100 s in total
    total_steps = args.steps + args.warmup
    sample = int(min(W.cpu_sample * 4, max(1 << 12, rate1 * cores * 0.6 * 100.0 / total_steps)))
    sample -= sample % cores
    if args.warmup:
        cpu_rate(kind, sample, args.warmup, cores)
    rate, wall = cpu_rate(kind, sample, args.steps, cores)
    desc = f"{sample} units/step x {args.steps} steps of the {cpu_port_desc(kind)}, {cores} processes"
    line = {
        "impl": "reference", "metric": W.metric, "value": rate, "unit": W.unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": W.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": W.name, "sample_units_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": W.unit, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": W.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json_line(line)


# ==================================================================================================
# GPU arm
# ==================================================================================================
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
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def measured_math_peaks():
    """FP32 / FP64 FMA and MUFU peaks measured on the box by tools/f2bench (SURVEY 8d asks for them beside hbm_gbs)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "math_peaks.json")))
    except Exception:
        return None


def measured_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json,
    written by scripts/make_traffic_json.py); None when there is no capture for this workload."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    from emei_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; emei_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    wl = WORKLOADS[args.workload](args, rank, world, dev)
    wl.setup()

    for i in range(W):
        wl.step(i)
    torch.cuda.synchronize()
    launches0 = _lib.launch_count
    graph = None
    if wl.use_graph:  # launch-bound steps: K launches captured once, replayed as one graph
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for i in range(K):
                    wl.step(i)
        launches = _lib.launch_count - launches0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # warm the instantiated graph / the step, and keep the GPU under load until nvidia-smi has delivered its first
    # samples (it needs a few hundred ms to start; a 4 ms C1 region would otherwise end before the first one)
    t_warm = time.perf_counter()
    while True:
        if graph is not None:
            graph.replay()
        else:
            wl.step(0)
        torch.cuda.synchronize()
        if rank != 0 or len(sampler.rows) >= 3 or time.perf_counter() - t_warm > 1.5:
            break
    _lib.call("emei_stats_reset", wl.stats.data_ptr(), torch.cuda.current_stream(dev).cuda_stream, launches=0)
    launches0 = _lib.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph is not None:
        # keep the device busy for ~0.1 ms while the host submits the graph, so that the events bracket the K steps
        # and not the host's graph-launch latency (~30 us: 15 % of a 20-step C2 region, nothing at K = 2000)
        torch.cuda._sleep(200_000)
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(K):
            wl.step(i)
    ev1.record()
    if world > 1:
        dist.all_reduce(wl.stats)  # end-of-rollout statistics: 2 doubles over NCCL/NVLink
    torch.cuda.synchronize()
    if graph is None:
        launches = _lib.launch_count - launches0
    if world > 1:
        dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    units = torch.tensor([float(wl.units)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units)
    ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = float(units.item()) * K / (ms_total * 1e-3)

    # ---------------- dominant-kernel duration for the roofline (CUDA events around that kernel alone)
    kern_ms = ms_per_step
    if not wl.use_graph and hasattr(wl, "env") and args.workload.startswith("c3"):
        from emei_b200 import engine

        kern_ms = engine.time_fused_scoring(wl.env, wl.obs, wl.pre, wl.act, reps=max(3, min(K, 10)))

    # ---------------- e2e: host inputs in, host results out, every step, through the public API
    e2e = None
    wl.setup_e2e()
    if wl.e2e_api is not None:
        e2e_steps = max(3, min(K, args.e2e_steps))
        e2e_steps = min(e2e_steps, getattr(wl, "e2e_max_steps", e2e_steps))
        for i in range(getattr(wl, "e2e_warmup", 2)):  # untimed: first-use work (step_host captures one CUDA graph per
            wl.step_e2e(i)                               # action buffer and ping-pong side the third time it sees the pair)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(e2e_steps):
            wl.step_e2e(i)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_units = float(units.item()) * getattr(wl, "e2e_units", wl.units) / wl.units
        e2e = {
            "value": e2e_units * e2e_steps / (float(te.item()) * 1e-3), "unit": wl.unit,
            "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h, "steps": e2e_steps, "api": wl.e2e_api,
        }
    clocks = sampler.finish() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src, sm_mhz = measured_peaks()
    achieved = wl.alg_bytes * wl.units / (kern_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
        "peak_source": peak_src, "algorithmic_bytes_per_unit": wl.alg_bytes, "kernel": wl.kernel,
        "kernel_ms_per_launch": kern_ms,
        "note": "duration = CUDA-event time of the dominant kernel per launch"
                + (" (timed region / launches, inter-launch gaps included)" if wl.use_graph else ""),
    }
    ipu = getattr(wl, "inst_per_unit", None)
    if wl.bound == "issue":  # no per-unit HBM traffic to speak of: the roofline is warp-instruction issue
        issue_peak = 148 * 4 * sm_mhz * 1e6
        ach = ipu * wl.units / 32.0 / (kern_ms * 1e-3)
        roofline = {
            "bound": "issue", "achieved": ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s", "frac": ach / issue_peak,
            "traffic": None, "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz), 1 warp instruction per scheduler per clock",
            "warp_inst_per_unit": ipu, "kernel": wl.kernel, "kernel_ms_per_launch": kern_ms,
            "hbm": {"algorithmic_bytes_per_unit": wl.alg_bytes, "achieved_gbs": achieved, "frac_of_hbm_peak": achieved / peak},
            "note": "duration = CUDA-event time per launch; instruction count per env-step from ncu smsp__inst_executed (profiles/)",
        }
    elif ipu:  # the second roofline of the north star: warp-instruction issue (148 SMs x 4 schedulers x clock)
        issue_peak = 148 * 4 * sm_mhz * 1e6
        t_math = ipu * wl.units / 32.0 / issue_peak
        t_hbm = wl.alg_bytes * wl.units / (peak * 1e9)
        roofline["math"] = {
            "warp_inst_per_unit": ipu, "issue_peak_warp_inst_per_s": issue_peak, "t_math_us": t_math * 1e6,
            "t_hbm_us": t_hbm * 1e6, "slower_bound": "math" if t_math > t_hbm else "hbm",
            "frac_of_slower_bound": max(t_math, t_hbm) / (kern_ms * 1e-3),
        }
    mp = measured_math_peaks()
    if mp is not None and ipu:
        tgt = roofline["math"] if "math" in roofline else roofline
        tgt["measured_math_peaks"] = {k: mp[k] for k in ("fp32_tflops", "fp64_tflops", "mufu_rcp_per_s", "mufu_sin_per_s") if k in mp}
        tgt["measured_math_peaks"]["source"] = "profiles/math_peaks.json (tools/f2bench on the B200 box)"
    tr = measured_traffic(args.workload)
    if tr is not None:
        roofline["traffic"] = tr["bytes_per_launch"]
        roofline["traffic_source"] = f"{tr['source']}: dram__bytes_read.sum + dram__bytes_write.sum of {tr['kernel'][:60]}... per launch (ncu --set full, cold L2; writes may still sit in the 126 MB L2 when the launch ends)"
        roofline["algorithmic_bytes_per_launch"] = wl.alg_bytes * wl.units
        if not wl.use_graph and wl.bound == "hbm":  # kernels much larger than L2: the capture's DRAM bytes at this run's duration
            roofline["dram_gbs_by_traffic"] = tr["bytes_per_launch"] / (kern_ms * 1e-3) / 1e9
            roofline["dram_frac_by_traffic"] = roofline["dram_gbs_by_traffic"] / peak
    cfg = {"workload": wl.name}
    cfg.update(wl.config())
    cfg["launch"] = (f"K={K} steps captured in one CUDA graph (programmatic dependent launches), replayed once" if wl.use_graph
                     else f"K={K} steps launched back to back on one stream") + "; CUDA events on the launching stream"
    cfg["parallelism"] = f"batch sharded over {world} rank(s), no data-path collective; NCCL all-reduce of the 2-double statistics at the end"
    line = {
        "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
        "dtype": wl.dtype, "data": "synthetic", "config": cfg, "roofline": roofline,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    # ---------------- CPU baseline on this box's host cores (bounded sample; oracle = the thing timed beside us)
    if world == 1 and not args.no_cpu:
        n_s = wl.cpu_sample
        # a bounded sample worth ~10 s of one core: calibrate with one pass, then size the pass count
        _, wall0 = cpu_rate(wl.cpu_kind, n_s, 1, 1)
        reps = int(max(2, min(200, round(10.0 / max(wall0, 1e-3)))))
        rate1, wall1 = cpu_rate(wl.cpu_kind, n_s, reps, 1)
        line["cpu_baseline"] = {
            "value": rate1, "unit": wl.unit, "cores": 1, "kind": "port",
            "sample": f"{n_s} units x {reps} passes of the {cpu_port_desc(wl.cpu_kind)}, {wall1:.1f} s",
        }
    emit_json_line(line)
    if world > 1:
        dist.destroy_process_group()


# The contract is ONE JSON line on stdout.  Libraries write to the process's stdout on their own (NCCL prints its
# version banner there with printf when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at
# stderr for the whole run and the JSON line alone goes to the real stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--ring", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--total-log2", type=int, default=0, help="c4 only: log2 of the total env count (default 26)")
    ap.add_argument("--horizon", type=int, default=None, help="rollout workloads: env-steps per launch (default 100; 32 with records)")
    args = ap.parse_args()
    args.horizon_set = args.horizon is not None
    if args.horizon is None:
        args.horizon = 100
    if args.steps is None:
        args.steps = 2000 if WORKLOADS[args.workload].use_graph else 20
        if args.impl == "reference":
            args.steps = 20
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
