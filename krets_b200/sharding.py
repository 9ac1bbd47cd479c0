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


def solve_sharded_chunked(total_lanes: int, solve_chunk: Callable[[int, int, "torch.Tensor"], None], *, chunk_lanes: int, row_shape, dtype,
                          device="cpu", gather_to: Optional[int] = None):
    """Chunk-pipelined sharded sweep.  This rank's contiguous lane block is solved chunk by chunk:
    `solve_chunk(lo, hi, out)` fills `out[0:hi-lo]` (shape [hi-lo, *row_shape]) with the results of the global lanes [lo, hi).

    gather_to=None (the default): results stay sharded on their GPUs — returns (local_lo, local_results); nothing crosses NVLink.
    gather_to=r: every chunk is sent to rank r as soon as it is solved (dist.isend: ncclSend on NCCL's stream / gloo on CPU) while the
    next chunk is being solved into the other buffer; the root posts the matching receives of one chunk round as ONE grouped
    operation (dist.batch_isend_irecv: ncclGroupStart/End) straight into the final [total_lanes, *row_shape] array, so the
    transfers of all ranks run concurrently and overlap the root's own chunk.  Returns (0, full) on the root, (local_lo, None) elsewhere.
    """
    import torch
    import torch.distributed as dist

    row_shape = tuple(row_shape)
    if not (dist.is_available() and dist.is_initialized()):
        world, rank = 1, 0
    else:
        world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(total_lanes, world, rank)

    def chunks(a, b):
        return [(c, min(c + chunk_lanes, b)) for c in range(a, b, chunk_lanes)]

    if gather_to is None or world == 1:
        local = torch.empty((hi - lo,) + row_shape, dtype=dtype, device=device)
        for c0, c1 in chunks(lo, hi):
            solve_chunk(c0, c1, local[c0 - lo:c1 - lo])
        return lo, local
    if rank == gather_to:
        full = torch.empty((total_lanes,) + row_shape, dtype=dtype, device=device)
        remote = {r: chunks(*shard_bounds(total_lanes, world, r)) for r in range(world) if r != rank}
        mine = chunks(lo, hi)
        rounds = max([len(mine)] + [len(v) for v in remote.values()])
        pending = []
        for k in range(rounds):
            ops = [dist.P2POp(dist.irecv, full[v[k][0]:v[k][1]], r) for r, v in remote.items() if k < len(v)]
            if ops:
                pending += dist.batch_isend_irecv(ops)
            if k < len(mine):
                solve_chunk(mine[k][0], mine[k][1], full[mine[k][0]:mine[k][1]])
        for req in pending:
            req.wait()
        return 0, full
    bufs = [torch.empty((min(chunk_lanes, max(hi - lo, 1)),) + row_shape, dtype=dtype, device=device) for _ in range(2)]
    reqs = [None, None]
    for k, (c0, c1) in enumerate(chunks(lo, hi)):
        b = k & 1
        if reqs[b] is not None:
            for req in reqs[b]:
                req.wait()                     # the send that last used this buffer has completed
        out = bufs[b][: c1 - c0]
        solve_chunk(c0, c1, out)
        reqs[b] = dist.batch_isend_irecv([dist.P2POp(dist.isend, out, gather_to)])
    for rq in reqs:
        if rq is not None:
            for req in rq:
                req.wait()
    return lo, None
