"""N > 1 host logic on CPU: world_size-2 gloo processes shard a lane batch, 'solve' their block with the numpy emulator of
the compiled lane program (test infrastructure standing in for the GPU) and gather; the result must equal the single-rank run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from krets_b200.sharding import shard_bounds, solve_sharded, solve_sharded_chunked


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 64, 65536, 1000001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker(rank, world, port, total, out_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import netgen
    from vm_emulator import run_chain
    import krets_b200 as kb
    from krets_b200 import api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    plan = kb.Plan(kb.parse_circuit_description(netgen.resistor_ladder(40)), api.MODE_DC, sweep_element=0)
    vals = np.arange(total) * 0.25 - 3.0

    def solve(lo, hi):
        X, _ = run_chain(plan, vals[None, lo:hi], False)
        return torch.from_numpy(np.ascontiguousarray(X.T))

    res = solve_sharded(total, solve, gather_to=0)
    if rank == 0:
        np.save(out_path, res.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [9, 16])
def test_two_rank_gloo_gather_matches_single_rank(tmp_path, total):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import netgen
    from vm_emulator import run_chain
    import krets_b200 as kb
    from krets_b200 import api
    out = str(tmp_path / "gathered.npy")
    port = 29500 + (os.getpid() % 2000) + total
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    got = np.load(out)
    plan = kb.Plan(kb.parse_circuit_description(netgen.resistor_ladder(40)), api.MODE_DC, sweep_element=0)
    vals = np.arange(total) * 0.25 - 3.0
    X, _ = run_chain(plan, vals[None, :], False)
    assert got.shape == (total, 41)
    assert np.array_equal(got, X.T)


def _worker_chunked(rank, world, port, total, chunk, out_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import netgen
    from vm_emulator import run_chain
    import krets_b200 as kb
    from krets_b200 import api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    plan = kb.Plan(kb.parse_circuit_description(netgen.resistor_ladder(40)), api.MODE_DC, sweep_element=0)
    vals = np.arange(total) * 0.25 - 3.0

    def solve(lo, hi, out):
        X, _ = run_chain(plan, vals[None, lo:hi], False)
        out.copy_(torch.from_numpy(np.ascontiguousarray(X.T)))

    lo, sharded = solve_sharded_chunked(total, solve, chunk_lanes=chunk, row_shape=(41,), dtype=torch.float64)          # default: stays sharded
    l2, full = solve_sharded_chunked(total, solve, chunk_lanes=chunk, row_shape=(41,), dtype=torch.float64, gather_to=0)
    np.save(f"{out_path}.shard{rank}.npy", np.concatenate([[lo], sharded.numpy().ravel()]))
    if rank == 0:
        np.save(out_path, full.numpy())
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("total,chunk", [(23, 4), (16, 16), (5, 1)])
def test_chunk_pipelined_gather_two_ranks(tmp_path, total, chunk):
    """world_size 2, gloo: results stay sharded by default; with gather_to the chunks travel while the next one is solved and land
    in lane order on the root (uneven shards, a last partial chunk, more rounds on one rank than on the other)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import netgen
    from vm_emulator import run_chain
    import krets_b200 as kb
    from krets_b200 import api
    out = str(tmp_path / "full.npy")
    port = 31500 + (os.getpid() % 2000) + total
    mp.spawn(_worker_chunked, args=(2, port, total, chunk, out), nprocs=2, join=True)
    plan = kb.Plan(kb.parse_circuit_description(netgen.resistor_ladder(40)), api.MODE_DC, sweep_element=0)
    vals = np.arange(total) * 0.25 - 3.0
    X, _ = run_chain(plan, vals[None, :], False)
    assert np.array_equal(np.load(out), X.T)
    for r in range(2):
        sh = np.load(f"{out}.shard{r}.npy")
        lo, hi = shard_bounds(total, 2, r)
        assert int(sh[0]) == lo and np.array_equal(sh[1:].reshape(hi - lo, 41), X.T[lo:hi])
