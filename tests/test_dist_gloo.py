"""world_size-2 gloo tests (CPU): the N>1 host logic -- contiguous sharding, the all-reduce of
[return_sum, done_count] and of the batch-wide sum of squared actions."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import emei_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emei_b200.dist import all_reduce_sum_, shard_range, world as world_fn

    assert world_fn() == (rank, world)
    n = 1001
    rng = np.random.default_rng(7)  # same global batch on every rank
    obs = rng.standard_normal((n, 18))
    pre = obs + 0.01 * rng.standard_normal((n, 18))
    act = rng.uniform(-1, 1, size=(n, 6))
    b, e = shard_range(n)
    # pass 1 on the local shard (what emei_sumsq_* computes on the device), then the exchange step
    local = torch.tensor([float(np.sum(np.square(act[b:e])))], dtype=torch.float64)
    glob = all_reduce_sum_(local.clone())
    p = O.HalfCheetahParams()
    r_local = O.halfcheetah_reward(obs[b:e], pre[b:e], act[b:e], p, sumsq=float(glob))
    r_ref = O.halfcheetah_reward(obs, pre, act, p)[b:e]
    ok_reward = bool(np.allclose(r_local, r_ref, rtol=1e-13, atol=1e-13))
    # end-of-rollout statistics
    stats = torch.tensor([float(r_local.sum()), float(e - b)], dtype=torch.float64)
    all_reduce_sum_(stats)
    ok_stats = abs(stats[0].item() - float(O.halfcheetah_reward(obs, pre, act, p).sum())) < 1e-8 and int(stats[1].item()) == n
    q.put((rank, b, e, ok_reward, ok_stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_scoring_and_stats():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, e0, okr0, oks0), (r1, b1, e1, okr1, oks1) = res
    assert (b0, e0, b1, e1) == (0, 501, 501, 1001)
    assert okr0 and okr1 and oks0 and oks1


def test_single_process_identity():
    from emei_b200.dist import all_reduce_sum_, is_distributed, shard_range, world

    assert not is_distributed() and world() == (0, 1)
    t = torch.tensor([1.5, 2.0], dtype=torch.float64)
    assert all_reduce_sum_(t) is t and t.tolist() == [1.5, 2.0]
    assert shard_range(10) == (0, 10)
