"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np

import oracle_py
from hybrid9_b200 import synth

STATE_FIELDS = ("h2osoi_liq", "zwt", "wa", "lai", "lai_litter", "plant_mass", "plant_foliage_mass",
                "plant_length", "rdepth", "rootr_col", "smp")


def make_oracle(world, nisurf=48, nyr=1, kind="f32", loop_order=1, smp_leak=0, nthreads=1):
    o = oracle_py.Oracle(kind)
    o.configure(world.nx, world.ny, nisurf, synth.ZI_DRIVER, nyr=nyr)
    o.set_soil(world.soil_tex, world.theta_s, world.hksat, world.bsw, world.psi_s, world.fmax)
    o.set_options(loop_order=loop_order, smp_leak=smp_leak, nthreads=nthreads)
    return o


THREAD_PER_CELL = 64   # h9_set_tuning block: thread-per-cell kernel, 64 threads per block, all
                       # registers: the small-shard step (all-deep / general straight-line tails)
THREAD_128REG = 1128   # thread-per-cell kernel compiled for <=128 registers: the throughput step
TWO_LANES = 4000       # the two-lanes-per-cell kernel (small shards)
FAST_KERNELS = [THREAD_PER_CELL, THREAD_128REG, TWO_LANES]
FAST_KERNEL_IDS = ["thread_per_cell", "thread_per_cell_128reg", "two_lanes"]


def make_gpu(world, nisurf=48, nyr=1, mode=0, device=0, block=0):
    """block: 0 = the library's automatic choice of the fast-mode stepping kernel (two lanes
    per cell up to ~9.5k cells, thread per cell above), THREAD_PER_CELL / TWO_LANES pin it."""
    from hybrid9_b200 import H9
    h = H9(device)
    h.configure(world.nx, world.ny, nisurf, synth.ZI_DRIVER, nyr=nyr)
    h.set_math(mode)
    if block:
        h.set_tuning(0, block)
    h.set_soil(world.soil_tex, world.theta_s, world.hksat, world.bsw, world.psi_s, world.fmax)
    return h


def day_slice(forcing, d):
    return {k: np.ascontiguousarray(v[d]) for k, v in forcing.items()}


def assert_state_equal(a, b, land, fields=STATE_FIELDS, what=""):
    for n in fields:
        x, y = getattr(a, n)[land], getattr(b, n)[land]
        same = (x == y) | (np.isnan(x) & np.isnan(y))
        assert same.all(), f"{what}{n}: {int((~same).sum())} of {same.size} values differ, " \
                           f"max abs {np.nanmax(np.abs(x.astype(np.float64) - y))}"


def assert_state_close(a, b, land, rtol, atol, fields=STATE_FIELDS, what="", smp_rtol=None):
    for n in fields:
        x = getattr(a, n)[land].astype(np.float64)
        y = getattr(b, n)[land].astype(np.float64)
        rt = smp_rtol if (n == "smp" and smp_rtol is not None) else rtol
        err = np.abs(x - y) - (atol + rt * np.abs(y))
        assert (err <= 0).all(), f"{what}{n}: max abs err {np.abs(x - y).max():.3e}, worst excess " \
                                 f"{err.max():.3e} at value {y.flat[np.argmax(err)]:.6g}"
