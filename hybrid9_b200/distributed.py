"""Multi-GPU plumbing: one process per GPU, land cells sharded in contiguous
latitude bands, NCCL only for what the driver gathers once per simulated year.

The reference exchanges nothing during time stepping (cells are independent:
HYDROLOGY.f90 only ever indexes the current (x,y)); its ranks meet again only in
the collective netCDF writers (WRITE_NET_CDF_3DR.f90:93-94,236-257).  So the
data-path has no collective at all; per simulated year this module does

  * an all-gather of the 13 annual-mean planes (HYBRID9.f90:263-291) so that the
    rank that writes axy<yyyy>.nc holds the whole grid, and
  * an FP64 all-reduce of the 8 budget partial sums of h9_annual_device.

On the GPUs both collectives run inside libh9gpu (h9_comm_init / h9_annual_collective,
include/h9gpu.h) on the ctx's own stream; torch.distributed only carries the 128-byte NCCL
id at start-up, as MPI_Bcast does for the Fortran host.  `gather_annual` is the same
exchange on torch tensors: the gloo tests use it to cover the host-side logic (padding,
ragged shards, scatter to the global grid) on CPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .host import partition_lat_bands
from .synth import World

ANNUAL_PLANES = 13
BUDGET_LEN = 8


def shard_world(world: World, rank: int, nranks: int):
    """Latitude band of `rank`: (sub-world, lat_s [1-based], lat_count, n_land of every rank)."""
    lat_s, lat_c, n_land = partition_lat_bands(world.soil_tex, world.theta_s, nranks)
    sub = world.window(1, int(lat_s[rank]), world.nx, int(lat_c[rank]))
    return sub, int(lat_s[rank]), int(lat_c[rank]), n_land


def shard_forcing(forcing: dict, lat_s: int, lat_c: int) -> dict:
    return {k: np.ascontiguousarray(v[:, lat_s - 1:lat_s - 1 + lat_c, :]) for k, v in forcing.items()}


class _CudaView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def device_tensor(ptr: int, shape, dtype=torch.float32, device=None) -> torch.Tensor:
    typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device=device)


def gather_annual(means: torch.Tensor, budget: torch.Tensor, n_land, group=None):
    """means: [13, stride_r] annual planes of this rank (first n_land[r] columns valid);
    budget: [8] float64 partial sums.  Returns (list of [13, n_land[r]] per rank on every
    rank, all-reduced budget).  Ranks may hold different cell counts: planes are padded to
    the largest shard for the all-gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nmax = int(max(int(x) for x in n_land))
    mine = torch.zeros((ANNUAL_PLANES, nmax), dtype=means.dtype, device=means.device)
    mine[:, :int(n_land[rank])] = means[:, :int(n_land[rank])]
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    total = budget.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return [p[:, :int(n_land[r])] for r, p in enumerate(parts)], total


def scatter_to_grid(parts, land_indices, lat_offsets, nx: int, ny: int, fill=np.nan):
    """Place every rank's compact planes on the global (ny, nx) grid: axy_* as
    WRITE_NET_CDF_3DR expects them.  land_indices[r] are block-local (y*nx + x)."""
    out = np.full((ANNUAL_PLANES, ny, nx), fill, dtype=np.float32)
    out[4] = 0.0  # axy_theta_total is zero-filled, INIT.f90:414
    flat = out.reshape(ANNUAL_PLANES, ny * nx)
    for r, p in enumerate(parts):
        idx = np.asarray(land_indices[r], dtype=np.int64) + (int(lat_offsets[r]) - 1) * nx
        flat[:, idx] = p if isinstance(p, np.ndarray) else p.detach().cpu().numpy()
    return out


def comm_init_from_torch(h9, group=None, device=None):
    """Bootstrap the library's own NCCL communicator (h9_comm_init) the way the Fortran host
    does with MPI_Bcast: rank 0 creates the id, the 128 bytes travel over the process group
    that already exists (NCCL or gloo), every rank joins."""
    from .host import COMM_ID_BYTES, comm_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    raw = comm_unique_id() if rank == 0 else bytes(COMM_ID_BYTES)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0, group=group)
    h9.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))
    return h9.comm_land_counts()


def h9_annual_collective(h9, iyr: int):
    """The per-year collective step for a live H9 context: one C-ABI call; the budget kernel,
    the FP64 all-reduce and the all-gather are stream-ordered behind the stepping kernel on the
    ctx's own stream (no host synchronisation, persistent buffers, one budget slot per year)."""
    h9.annual_collective(iyr)


def fetch_gathered(h9, iyr: int, n_land):
    """Host copies of what h9_annual_collective left on the device: per-rank compact planes
    [13, n_land[r]] and the all-reduced budget of year iyr."""
    parts = [h9.get_gathered_annual(r, int(n)) for r, n in enumerate(n_land)]
    return parts, h9.get_budget(iyr)
