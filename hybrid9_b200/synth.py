"""Seeded synthetic world and PGF-shaped forcing.

None of the reference's input data exists here (PGF forcing, BNU soils, HWSD
textures under /scratch/adf10/...), so the tests and the benchmark use a
deterministic stand-in of the same shapes, units and value ranges:

* grid nx x ny (720x360 at 0.5 deg; 1440x720 at 0.25 deg), row 0 = northernmost
  (INIT.f90:145), column 0 = westernmost (INIT.f90:142);
* `soil_tex` int32 in 0..13 with exactly `n_land` cells passing the land
  predicate of HYBRID9.f90:122-123, plus cells of class 13 and land-textured
  cells with theta_s == 0 that must be rejected by it;
* per-layer theta_s / hksat / bsw / psi_s in the units INIT.f90:610-628 produces;
* daily forcing (ndays, ny, nx) for the seven PGF fields of READ_PGF.f90.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

F32 = np.float32
ZI_DRIVER = np.array([0.0, 45.0, 91.0, 166.0, 289.0, 493.0, 829.0, 1383.0, 2296.0, 5000.0],
                     dtype=F32)  # EXECUTE/driver.txt:17-26
NISURF_DRIVER = 48               # EXECUTE/driver.txt:2
N_LAND_HALF_DEG = 67420
N_LAND_QUARTER_DEG = 269680
# BASELINE.json config 2: a 100x74 regional block (lon_s, lat_s, lon_c, lat_c) of the seed-9
# 0.5 deg world holding exactly 2,500 land cells
REGIONAL_WINDOW = (73, 8, 100, 74)


@dataclass
class World:
    nx: int
    ny: int
    soil_tex: np.ndarray   # (ny, nx) int32
    theta_s: np.ndarray    # (ny, nx, 8)
    hksat: np.ndarray      # (ny, nx, 8) mm/s
    bsw: np.ndarray        # (ny, nx, 8)
    psi_s: np.ndarray      # (ny, nx, 8) mm (negative)
    fmax: np.ndarray       # (ny, nx)
    lat: np.ndarray        # (ny,) degrees
    lon: np.ndarray        # (nx,) degrees
    seed: int

    @property
    def land(self) -> np.ndarray:
        from .state import land_mask
        return land_mask(self.soil_tex, self.theta_s)

    def window(self, lon_s: int, lat_s: int, lon_c: int, lat_c: int) -> "World":
        """Block (lon_s, lat_s are 1-based like CONTROL.f90:49-50)."""
        ys, xs = slice(lat_s - 1, lat_s - 1 + lat_c), slice(lon_s - 1, lon_s - 1 + lon_c)
        c = np.ascontiguousarray
        return World(lon_c, lat_c, c(self.soil_tex[ys, xs]), c(self.theta_s[ys, xs]),
                     c(self.hksat[ys, xs]), c(self.bsw[ys, xs]), c(self.psi_s[ys, xs]),
                     c(self.fmax[ys, xs]), self.lat[ys].copy(), self.lon[xs].copy(), self.seed)


def make_world(nx: int = 720, ny: int = 360, n_land: int | None = None, seed: int = 9,
               n_class13: int | None = None, n_zero_theta: int | None = None,
               n_missing_lambda: int = 0) -> World:
    rng = np.random.default_rng(seed)
    res = 360.0 / nx
    lon = (-180.0 + res / 2 + res * np.arange(nx)).astype(np.float64)
    lat = (90.0 - res / 2 - res * np.arange(ny)).astype(np.float64)
    if n_land is None:
        n_land = int(round(N_LAND_HALF_DEG * (nx * ny) / (720.0 * 360.0)))
    if n_class13 is None:
        n_class13 = min(200, max(1, n_land // 300))
    if n_zero_theta is None:
        n_zero_theta = min(200, max(1, n_land // 300))
    # smooth deterministic "continent" field, land restricted to -56 <= lat <= 84
    lo, la = np.meshgrid(np.deg2rad(lon), np.deg2rad(lat))
    f = (np.sin(2 * lo + 0.7) * np.cos(1.5 * la) + 0.6 * np.sin(3 * lo - 1.1) * np.sin(2 * la + 0.4)
         + 0.5 * np.cos(lo + 2.0) + 0.35 * np.sin(5 * lo + 3 * la))
    ok = (lat[:, None] >= -56.0) & (lat[:, None] <= 84.0) & np.ones((1, nx), bool)
    f = np.where(ok, f, -np.inf)
    n_cand = n_land + n_class13 + n_zero_theta
    if n_cand > int(ok.sum()):
        raise ValueError("n_land too large for the grid")
    flat = np.argsort(-f, axis=None, kind="stable")[:n_cand]
    soil_tex = np.zeros(ny * nx, np.int32)
    soil_tex[flat] = rng.integers(1, 13, size=n_cand).astype(np.int32)
    special = rng.permutation(flat)
    cls13, zero_th = special[:n_class13], special[n_class13:n_class13 + n_zero_theta]
    soil_tex[cls13] = 13
    soil_tex = soil_tex.reshape(ny, nx)

    shp = (ny, nx)
    # vertically correlated layers: a per-cell base plus a small per-layer perturbation
    def layered(lo_, hi_, spread):
        base = rng.uniform(lo_, hi_, size=shp)
        out = base[..., None] + rng.uniform(-spread, spread, size=shp + (8,))
        return np.clip(out, lo_, hi_)

    theta_s = layered(0.30, 0.55, 0.03).astype(F32)
    ks_cm_day = np.exp(layered(np.log(0.5), np.log(488.0), 0.3))          # notes.txt:187
    hksat = (F32(10.0) * ks_cm_day.astype(F32) / F32(86400.0)).astype(F32)  # INIT.f90:614
    lam = layered(0.08, 0.35, 0.02).astype(F32)
    bsw = (F32(1.0) / lam).astype(F32)                                    # INIT.f90:628
    psi_s = (-layered(5.0, 80.0, 5.0) * 10.0).astype(F32)                 # INIT.f90:616 (cm -> mm)
    fmax = rng.uniform(0.1, 0.6, size=shp).astype(F32)                    # INIT.f90:673
    theta_s.reshape(-1, 8)[zero_th] = 0.0
    if n_missing_lambda:
        land_flat = np.setdiff1d(flat, np.concatenate([cls13, zero_th]))
        pick = rng.choice(land_flat, size=n_missing_lambda, replace=False)
        bsw.reshape(-1, 8)[pick] = F32(1.0) / F32(1.0e-8)                 # INIT.f90:624-628
    w = World(nx, ny, soil_tex, theta_s, hksat, bsw, psi_s, fmax, lat, lon, seed)
    return w


def make_forcing(world: World, ndays: int, seed: int = 9, doy0: int = 1,
                 out: dict | None = None, land_only: bool = True) -> dict:
    """Seven PGF-shaped daily fields, each (ndays, ny, nx) float32.

    `out` may hold preallocated (e.g. pinned) arrays.  Ocean cells are 0 when
    land_only (the reference never reads them, HYBRID9.f90:122)."""
    rng = np.random.default_rng(seed + 1000003)
    ny, nx = world.ny, world.nx
    names = ("tas", "rlds", "rsds", "huss", "ps", "pr", "rhs")
    if out is None:
        out = {k: np.zeros((ndays, ny, nx), F32) for k in names}
    else:
        for k in names:
            out[k][...] = 0
    mask = world.land if land_only else np.ones((ny, nx), bool)
    yy, xx = np.nonzero(mask)
    n = yy.size
    lat = np.deg2rad(world.lat[yy])
    ps0 = rng.uniform(60e3, 103e3, size=n)
    chunk = max(1, min(ndays, int(4e6 // max(n, 1)) or 1))
    for d0 in range(0, ndays, chunk):
        nd = min(chunk, ndays - d0)
        doy = ((doy0 - 1 + d0 + np.arange(nd)) % 365 + 1)[:, None].astype(np.float64)
        tas = (288.0 - 45.0 * np.sin(lat) ** 2
               + 15.0 * np.cos(2 * np.pi * (doy - 200.0) / 365.0) * np.sin(lat)
               + rng.normal(0.0, 3.0, size=(nd, n)))
        tas = np.clip(tas, 220.0, 320.0)
        decl = np.deg2rad(23.44) * np.sin(2 * np.pi * (doy - 80.0) / 365.0)
        rsds = np.maximum(0.0, 340.0 * np.cos(lat - decl)) * rng.uniform(0.3, 1.0, size=(nd, n))
        rlds = 0.8 * 5.67e-8 * tas ** 4 + rng.normal(0.0, 15.0, size=(nd, n))
        ps = ps0 + rng.normal(0.0, 500.0, size=(nd, n))
        rhs = np.clip(rng.normal(65.0, 20.0, size=(nd, n)), 5.0, 100.0)
        tc = tas - 273.16
        esat_pa = 610.8 * np.exp(17.27 * tc / (tc + 237.3))
        huss = 0.622 * (rhs / 100.0) * esat_pa / ps
        wet = rng.random(size=(nd, n)) < 0.3
        pr = np.where(wet, rng.exponential(8e-5, size=(nd, n)), 0.0)
        for k, v in (("tas", tas), ("rlds", rlds), ("rsds", rsds), ("huss", huss), ("ps", ps),
                     ("pr", pr), ("rhs", rhs)):
            out[k][d0:d0 + nd, yy, xx] = v.astype(F32)
    return out


def compact_world(world: World, n: int | None = None, start: int = 0) -> World:
    """The land cells [start, start+n) of `world` (reference iteration order) as a
    1-row block (lon_c = n, lat_c = 1): a bounded sample for CPU-side runs."""
    yy, xx = np.nonzero(world.land)
    if n is None:
        n = yy.size - start
    yy, xx = yy[start:start + n], xx[start:start + n]
    c = lambda a: np.ascontiguousarray(a[yy, xx][None])  # noqa: E731
    return World(int(yy.size), 1, c(world.soil_tex), c(world.theta_s), c(world.hksat),
                 c(world.bsw), c(world.psi_s), c(world.fmax), world.lat[yy[:1]].copy(),
                 world.lon[xx].copy(), world.seed)


def compact_forcing(world: World, forcing: dict, n: int | None = None, start: int = 0,
                    ndays: int | None = None) -> dict:
    """Forcing of the same cells as compact_world(world, n, start): (ndays, 1, n)."""
    yy, xx = np.nonzero(world.land)
    if n is None:
        n = yy.size - start
    yy, xx = yy[start:start + n], xx[start:start + n]
    nd = forcing["tas"].shape[0] if ndays is None else ndays
    return {k: np.ascontiguousarray(v[:nd, yy, xx][:, None, :]) for k, v in forcing.items()}


def randomize_state(world: World, state, seed: int = 11):
    """Branch-coverage state: water table anywhere in 0..12 m
    (jwt = 0..8), layers from nearly dry to over-saturated, aquifer up to its cap,
    LAI from the floor to a closed canopy."""
    from .state import geometry
    rng = np.random.default_rng(seed)
    land = world.land
    _, dz, _ = geometry(ZI_DRIVER, 48)
    shp = land.shape
    st = state.copy()
    frac = rng.uniform(0.005, 1.1, size=shp + (8,))
    st.h2osoi_liq[...] = np.where(land[..., None], frac * world.theta_s * dz[1:9], 0).astype(F32)
    # half the cells: water table inside a uniformly chosen soil layer (jwt = 0..7 all
    # populated); the rest below the column (jwt = 8), down to 12 m
    zi_m = ZI_DRIVER.astype(np.float64) / 1000.0
    lay = rng.integers(0, 8, size=shp)
    inside = zi_m[lay] + rng.uniform(0.0, 1.0, size=shp) * (zi_m[lay + 1] - zi_m[lay])
    below = rng.uniform(zi_m[8], 12.0, size=shp)
    zwt = np.where(rng.random(size=shp) < 0.5, inside, below)
    st.zwt[...] = np.where(land, zwt, 0).astype(F32)
    st.wa[...] = np.where(land, rng.uniform(3000.0, 5100.0, size=shp), 0).astype(F32)
    lai = np.where(rng.random(size=shp) < 0.25, 0.001, rng.uniform(0.1, 7.0, size=shp))
    st.lai[...] = np.where(land, lai, 0).astype(F32)
    st.lai_litter[...] = np.where(land, rng.uniform(0.0005, 0.5, size=shp), 0).astype(F32)
    pm = rng.uniform(0.5, 3000.0, size=shp)
    st.plant_mass[..., 0] = np.where(land, pm, 0).astype(F32)
    st.plant_foliage_mass[..., 0] = np.where(land, pm * rng.uniform(0.01, 0.1, size=shp), 0).astype(F32)
    smp = -np.exp(rng.uniform(np.log(50.0), np.log(3.0e5), size=shp + (8,)))
    st.smp[...] = np.where(land[..., None], smp, 0).astype(F32)
    # a root profile with mass in deeper layers, normalised like GROW leaves it
    r = rng.dirichlet(np.ones(8) * 0.7, size=shp)
    st.rootr_col[..., :8] = np.where(land[..., None], r, 0).astype(F32)
    st.rootr_col[..., 8] = 0
    return st
