"""Restart files and a spin-up controller.

The reference has neither: every run re-initialises from INIT.f90:707-811 and "use restarts as
well" is a TODO (`notes.txt:12`).  Both are host-side conveniences above the C ABI; the physics
stays in libh9gpu.so.

* `save_restart` / `load_restart`: the arrays of `h9_get_state` / `h9_set_state` (module SHARED's
  per-cell state plus `smp`) in one flat little-endian file: a 64-byte
  header (magic, version, lat_c, lon_c) followed by the arrays in `H9State` field order, each in
  the reference's memory order.  A Fortran host reads it with one stream-access READ per array.
  A restart taken on a year boundary continues bit for bit (tests/test_gpu_restart.py).
* `spin_up`: cycles a block of forcing years until the block-mean total soil water
  (`axy_theta_total`, HYBRID9.f90:290) of the last year drifts by less than `tol_mm` per cycle,
  or `max_cycles` is reached -- config 4 of BASELINE.json (multi-decade spin-up).
"""
from __future__ import annotations

import struct

import numpy as np

from .state import H9State

MAGIC = b"H9RESTART\0\0\0"
VERSION = 1


def save_restart(path: str, st: H9State) -> None:
    lat_c, lon_c = st.zwt.shape
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<iii", VERSION, lat_c, lon_c) + b"\0" * (64 - len(MAGIC) - 12))
        for n in st.names():
            a = getattr(st, n)
            dt = "<i4" if n == "nplants" else "<f4"
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())


def load_restart(path: str) -> H9State:
    with open(path, "rb") as f:
        head = f.read(64)
        if head[:len(MAGIC)] != MAGIC:
            raise ValueError(f"{path}: not an h9 restart file")
        version, lat_c, lon_c = struct.unpack("<iii", head[len(MAGIC):len(MAGIC) + 12])
        if version != VERSION:
            raise ValueError(f"{path}: restart version {version}, expected {VERSION}")
        st = H9State.zeros(lat_c, lon_c)
        for n in st.names():
            a = getattr(st, n)
            dt = "<i4" if n == "nplants" else "<f4"
            raw = f.read(a.size * 4)
            if len(raw) != a.size * 4:
                raise ValueError(f"{path}: truncated at '{n}'")
            setattr(st, n, np.frombuffer(raw, dtype=dt).reshape(a.shape).astype(a.dtype).copy())
        if f.read(1):
            raise ValueError(f"{path}: trailing bytes")
    return st


def spin_up(h, year_index, forcing, land, max_cycles: int = 50, tol_mm: float = 0.5):
    """Repeat `h.run_days(year_index, forcing)` (one or more whole years; year indices 1..n) until
    the land-mean of the last year's `theta_total` changes by less than `tol_mm` between two
    consecutive cycles.  Returns a list of (cycle, land-mean theta_total in mm, drift in mm)."""
    year_index = np.ascontiguousarray(year_index, np.int32)
    last = int(year_index[-1])
    hist, prev = [], None
    for cycle in range(1, max_cycles + 1):
        rc = h.run_days(year_index, forcing)
        if rc != 0:
            raise RuntimeError(f"physics fault {rc} in spin-up cycle {cycle}: {h.get_fault()}")
        tot = float(np.mean(h.get_annual(last)["theta_total"][land], dtype=np.float64))
        drift = float("inf") if prev is None else abs(tot - prev)
        hist.append((cycle, tot, drift))
        if drift < tol_mm:
            break
        prev = tot
    return hist
