"""Multi-GPU lane sharding: lanes are independent, so each rank solves a contiguous block of lanes on its
own GPU with zero inter-GPU traffic; the only collective is the final gather of result shards (NCCL over NVLink on GPUs,
gloo in the CPU tests).  One process per GPU, launched with torchrun."""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(total_lanes: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous lane block [lo, hi) of `rank`: [r*L/G, (r+1)*L/G) — remainders spread over the first ranks' boundaries."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    lo = (total_lanes * rank) // world_size
    hi = (total_lanes * (rank + 1)) // world_size
    return lo, hi


def solve_sharded(total_lanes: int, solve_shard: Callable[[int, int], "torch.Tensor"], *, gather_to: Optional[int] = 0):
    """Run `solve_shard(lo, hi)` for this rank's lane block and gather the [lanes_r, ...] result shards in lane order.

    With torch.distributed uninitialised this is a single-rank call.  gather_to=None skips the collective (results stay
    sharded on their GPUs); otherwise rank `gather_to` returns the concatenated result and the others return their shard.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return solve_shard(0, total_lanes)
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(total_lanes, world, rank)
    local = solve_shard(lo, hi)
    if gather_to is None:
        return local
    # shards may differ by one lane: gather sizes first, then pad to the largest shard
    sizes = [shard_bounds(total_lanes, world, r)[1] - shard_bounds(total_lanes, world, r)[0] for r in range(world)]
    biggest = max(sizes)
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local_p = torch.cat([local, pad], dim=0)
    else:
        local_p = local
    bufs = [torch.empty_like(local_p) for _ in range(world)] if rank == gather_to else None
    dist.gather(local_p, bufs, dst=gather_to)
    if rank != gather_to:
        return local
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
