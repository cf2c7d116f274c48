"""Batch sharding across the GPUs of one box.

Instances are independent, so the batch is cut into contiguous slices [rank*B/G, (rank+1)*B/G) -- one
process per GPU, no collective on the solve path -- and the per-instance results (T*, J*, status: <= 16 B
per solve) are exchanged by ONE final all_gather.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)
is plumbing only."""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition: the first B % world ranks get one extra instance."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(B: int, world: int):
    return [shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world)]


def all_gather_ragged(local: torch.Tensor, sizes: Sequence[int]) -> torch.Tensor:
    """all_gather of per-rank slices whose leading dimension differs by at most one (padded exchange)."""
    world = dist.get_world_size()
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def sharded_select(select_fn: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], x0_all: torch.Tensor):
    """Run `select_fn` (x0_slice -> (T*, J*, status) tensors) on this rank's slice of x0_all [B, n] and
    return the gathered (T*, J*, status) for the whole batch on every rank.  Works unsharded too."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return select_fn(x0_all)
    world, rank = dist.get_world_size(), dist.get_rank()
    B = x0_all.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    T, J, st = select_fn(x0_all[lo:hi])
    sizes = shard_sizes(B, world)
    return all_gather_ragged(T, sizes), all_gather_ragged(J, sizes), all_gather_ragged(st, sizes)
