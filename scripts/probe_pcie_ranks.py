"""Concurrent pinned-copy ceiling of the box: every rank copies at the same time (torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/probe_pcie_ranks.py

Prints, for rank counts 1 (rank 0 alone) and N (all together): per-GPU and aggregate D2H / H2D / bidirectional
bandwidth with 64 MiB pinned buffers, plus each GPU's PCI bus id, NUMA node and the CPU affinity of its rank --
what bounds the end-to-end (host-buffer) throughput of step_host when N ranks share one host."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import datetime

    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

from emei_b200.dist import bind_to_gpu_numa, gpu_numa_info  # noqa: E402

info = gpu_numa_info(local)
bound = bind_to_gpu_numa(local) if os.environ.get("EMEI_BIND_NUMA", "1") == "1" else None
nbytes = 64 << 20
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream(dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def d2h():
    h.copy_(d, non_blocking=True)


def h2d():
    d2.copy_(h2, non_blocking=True)


def both():
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
    h.copy_(d, non_blocking=True)


def report(tag, active):
    out = []
    for name, fn, mult in (("D2H", d2h, 1), ("H2D", h2d, 1), ("both", both, 2)):
        if active:
            t = timed(fn)
            gbs = mult * nbytes / t / 1e9
        else:
            if world > 1:
                dist.barrier()
            gbs = 0.0
        v = torch.tensor([gbs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v)
        out.append((name, gbs, float(v.item())))
        if world > 1:
            dist.barrier()
    if rank == 0:
        print(f"{tag}: " + "  ".join(f"{n} {mine:6.1f} GB/s on rank 0, {tot:7.1f} GB/s aggregate" for n, mine, tot in out), flush=True)


gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, (rank, info, bound, sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0))))
else:
    gathered = [(rank, info, bound, sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0)))]
if rank == 0:
    for g in gathered:
        print(f"rank {g[0]}: gpu {g[1]}  bound -> {g[2]}  affinity {g[4]} cpus (first {g[3]})", flush=True)
report("1 rank copying ", rank == 0)
report(f"{world} ranks copying", True)
if world > 1:
    dist.destroy_process_group()
