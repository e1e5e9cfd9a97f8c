"""Multi-GPU plumbing: one process per GPU (torchrun), the env/transition batch sharded in
contiguous slices, and the only two exchanges the path has:

  1. end-of-rollout statistics: all-reduce(SUM) of [return_sum, done_count] (2 doubles);
  2. the batch-wide sum of squared actions of Hopper/HalfCheetah's control cost
     (hopper.py:98, half_cheetah.py:61): all-reduce(SUM) of 1 double between the two passes.

Payloads are 8-16 bytes (latency-bound, NCCL over NVLink/NVSwitch); the data path itself has no
collective.  With ``torch.distributed`` uninitialised every function is the single-process identity.
Works with the ``gloo`` backend on CPU tensors too (used by the world_size-2 CPU tests).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> Tuple[int, int]:
    """(rank, world_size)."""
    if is_distributed():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of a batch of ``total`` units owned by ``rank``.
    Remainder units go to the lowest ranks; global env ids (begin + local index) key the Philox
    streams, so sampled initial states do not depend on the world size."""
    if rank is None or world_size is None:
        r, w = world()
        rank = r if rank is None else rank
        world_size = w if world_size is None else world_size
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(int(total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce (identity when not distributed).  Stays on the tensor's device and
    stream: no host synchronisation."""
    if is_distributed() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


# ---------------------------------------------------------------------------------------------------------
# host placement of a rank: the end-to-end (host-buffer) path moves 4 + 21 bytes per env-step over PCIe, so where
# the pinned staging pages live matters when several ranks share one host
# ---------------------------------------------------------------------------------------------------------
def gpu_numa_info(device_index: int) -> dict:
    """PCI bus id, NUMA node and local CPU list of a GPU, read from sysfs (None where the kernel does not say)."""
    import os

    out = {"pci": None, "numa_node": None, "local_cpulist": None}
    try:
        pci = torch.cuda.get_device_properties(device_index).pci_bus_id  # torch >= 2.5
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        bdf = f"{dom:04x}:{pci:02x}:{dev:02x}.0"
    except Exception:
        return out
    out["pci"] = bdf
    base = f"/sys/bus/pci/devices/{bdf}"
    for key, name in (("numa_node", "numa_node"), ("local_cpulist", "local_cpulist")):
        try:
            out[key] = open(os.path.join(base, name)).read().strip()
        except OSError:
            pass
    return out


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa(device_index: int):
    """Pin the calling process to the CPUs local to its GPU's NUMA node BEFORE it allocates pinned staging memory
    (first-touch places those pages on that node, next to the GPU's PCIe root port).  Returns the CPU set it bound
    to, or None when the host exposes no usable topology (one NUMA node, a container without sysfs, a cpuset that
    excludes the local CPUs) -- in which case nothing is changed."""
    import os

    info = gpu_numa_info(device_index)
    txt = info.get("local_cpulist")
    if not txt:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        local = _parse_cpulist(txt) & allowed
        if not local or local == allowed:
            return None
        os.sched_setaffinity(0, local)
        return sorted(local)
    except (OSError, ValueError):
        return None
