"""world_size-2 (and 3) gloo runs of the only exchange step the path has: the per-year
all-gather of the annual-mean planes and the FP64 budget all-reduce, plus the
latitude-band sharding of the host arrays."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hybrid9_b200 import distributed as h9d
from hybrid9_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        w = synth.make_world(nx=72, ny=36, seed=9)
        sub, lat_s, lat_c, n_land = h9d.shard_world(w, rank, world_size)
        nc = int(sub.land.sum())
        assert nc == int(n_land[rank])
        stride = (nc + 127) // 128 * 128
        # fake annual planes: plane p of global cell g holds p*1e6 + g
        gidx = np.flatnonzero(sub.land.ravel()) + (lat_s - 1) * w.nx
        means = torch.full((13, stride), -1.0, dtype=torch.float32)
        means[:, :nc] = torch.tensor(np.arange(13)[:, None] * 1.0e6 + gidx[None, :], dtype=torch.float32)
        budget = torch.tensor([float(nc), 1.0, 0, 0, 0, float(nc), 0, 0], dtype=torch.float64)
        parts, total = h9d.gather_annual(means, budget, n_land)
        land_idx = [None] * world_size
        offs = [None] * world_size
        for r in range(world_size):
            s_r, ls, lc, _ = h9d.shard_world(w, r, world_size)
            land_idx[r] = np.flatnonzero(s_r.land.ravel())
            offs[r] = ls
        grid = h9d.scatter_to_grid(parts, land_idx, offs, w.nx, w.ny)
        q.put((rank, grid, total.numpy(), int(w.land.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_annual_gather_and_budget_allreduce(world_size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world_size)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synth.make_world(nx=72, ny=36, seed=9)
    land = w.land
    g = np.arange(w.ny * w.nx).reshape(w.ny, w.nx)
    for rank, grid, total, nland in res:
        assert nland == land.sum()
        assert total[0] == land.sum() and total[1] == world_size and total[5] == land.sum()
        for p in range(13):
            assert np.array_equal(grid[p][land], (p * 1.0e6 + g[land]).astype(np.float32))
        assert np.isnan(grid[0][~land]).all() and np.all(grid[4][~land] == 0)


def test_shards_tile_the_block_in_reference_order():
    w = synth.make_world(nx=72, ny=36, seed=9)
    f = synth.make_forcing(w, 2, seed=1)
    for n in (1, 2, 4, 8):
        seen = []
        for r in range(n):
            sub, ls, lc, n_land = h9d.shard_world(w, r, n)
            seen.append(np.flatnonzero(sub.land.ravel()) + (ls - 1) * w.nx)
            fs = h9d.shard_forcing(f, ls, lc)
            assert fs["tas"].shape == (2, lc, w.nx)
            assert np.array_equal(fs["tas"], f["tas"][:, ls - 1:ls - 1 + lc])
        # concatenation of the shards' compact orders == the unsharded compact order
        assert np.array_equal(np.concatenate(seen), np.flatnonzero(w.land.ravel()))
