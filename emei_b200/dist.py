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
