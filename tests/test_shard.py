"""Tile sharding (host logic) incl. a world_size-2 gloo run on CPU: each rank computes its own
tiles with no data exchange, only control-plane reductions (max time, checksum sum)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from gfx_ocean_b200.shard import checksum, fan_out, rank_of_tile, tiles_of_rank


@pytest.mark.parametrize("world,n", [(1, 8), (2, 8), (4, 64), (8, 64), (3, 8), (8, 5), (2, 1)])
def test_partition_is_contiguous_and_complete(world, n):
    parts = [tiles_of_rank(r, world, n) for r in range(world)]
    assert sum(parts, []) == list(range(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1
    for t in range(n):
        assert t in parts[rank_of_tile(t, world, n)]


def test_partition_rejects_bad_arguments():
    with pytest.raises(ValueError):
        tiles_of_rank(2, 2, 8)
    with pytest.raises(ValueError):
        rank_of_tile(9, 2, 8)


def _worker(rank, world, port, n_tiles, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gfx_ocean_b200.spectrum import synthetic_tile
    from oracle.ocean_oracle import COracle
    o = COracle()
    o.set_num_threads(1)
    sums = torch.zeros(n_tiles, dtype=torch.float64)
    # launch fan-out: only rank 0 knows which frames to run; the others learn it from the broadcast block
    first, count, dt = fan_out(*((4, 1, 0.25) if rank == 0 else (-1, -1, -1.0)))
    assert (first, count, dt) == (4, 1, 0.25)
    for g in tiles_of_rank(rank, world, n_tiles):          # compute stands in for the GPU frame
        h0, w = synthetic_tile(n, g)
        sums[g] = checksum(o.frame(h0, w, first * dt, n, prec="f64"))
    elapsed = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)             # control plane only
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((sums.tolist(), float(elapsed)))
    dist.destroy_process_group()


def test_fan_out_is_the_identity_without_a_process_group():
    assert fan_out(3, 200, 0.016) == (3, 200, 0.016)


def test_two_rank_gloo_run_matches_single_rank():
    n_tiles, n = 4, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_tiles, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums, elapsed = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from gfx_ocean_b200.spectrum import synthetic_tile
    from oracle.ocean_oracle import COracle
    o = COracle()
    want = [checksum(o.frame(*synthetic_tile(n, g), 1.0, n, prec="f64")) for g in range(n_tiles)]
    np.testing.assert_allclose(sums, want, rtol=1e-12)
    assert elapsed == 2.0                                    # max over ranks
