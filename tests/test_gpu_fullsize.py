"""Size-independent properties at BASELINE.json's full size (0.5 deg global land mask,
67,420 cells), where the oracle would take minutes: the reference's in-code invariants,
determinism, and the sharding property behind the multi-GPU layout (cells are
independent, so any latitude-band split reproduces the unsplit run bit for bit)."""
import numpy as np
import pytest

from helpers import THREAD_PER_CELL, assert_state_equal, make_gpu
from hybrid9_b200 import MATH_FAST, synth
from hybrid9_b200.host import partition_lat_bands
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu
ND = 6


@pytest.fixture(scope="module")
def big():
    w = synth.make_world()  # 720 x 360, exactly 67,420 land cells
    f = synth.make_forcing(w, ND, seed=9)
    return w, f


def run(w, f, st=None, block=0):
    h = make_gpu(w, mode=MATH_FAST, block=block)
    h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER) if st is None else st)
    rc = h.run_days(np.ones(f["tas"].shape[0], np.int32), f)
    out = (rc, h.get_state(), h.get_annual(1), h.get_fault())
    h.close()
    return out


def test_full_size_invariants_and_determinism(big):
    w, f = big
    assert int(w.land.sum()) == synth.N_LAND_HALF_DEG
    rc, st, ann, fault = run(w, f)
    land = w.land
    assert rc == 0 and fault.n_faulted == 0          # |w1-w0| <= 0.1 mm at every one of 19.4 M cell-steps
    assert (st.zwt[land] >= 0).all() and (st.zwt[land] <= 80).all()   # HYDROLOGY.f90:1122-1123
    assert (st.wa[land] <= 5000).all()                                # :1054
    assert (st.h2osoi_liq[land] >= 0.01 * (1 - 1e-5)).all()           # :1161-1205
    dz = np.array([45, 46, 75, 123, 204, 336, 554, 913], np.float32)
    cap = np.maximum(np.float32(0.01), w.theta_s) * dz
    assert (st.h2osoi_liq[land][:, 1:] <= cap[land][:, 1:] * (1 + 1e-5)).all()  # :1131-1137
    assert (st.lai[land] >= np.float32(0.001)).all()                  # GROW.f90:163
    assert np.isfinite(st.smp[land]).all() and (st.smp[land] >= -1e8).all()     # smpmin
    assert np.all(ann["evap"][land] == 0) and np.isfinite(ann["theta"][land]).all()
    assert np.isnan(ann["npp"][~land]).all()
    rc2, st2, ann2, _ = run(w, f)
    assert_state_equal(st, st2, land)
    for k in ann:
        assert np.array_equal(ann[k], ann2[k], equal_nan=True)


@pytest.mark.parametrize("nranks", [2, 8])
def test_latitude_band_shards_reproduce_the_whole(big, nranks):
    """Bit for bit, with the shards' stepping kernel pinned to thread-per-cell (h9_set_tuning):
    every launch shape and both builds of that kernel give the same bits (the whole grid steps
    with the 128-register build, the shards with the all-register build and its straight-line
    tails), so any split reproduces the unsplit run."""
    w, f = big
    _, whole, ann, _ = run(w, f)  # automatic: the 128-register build (throughput step) at this size
    lat_s, lat_c, n_land = partition_lat_bands(w.soil_tex, w.theta_s, nranks)
    assert n_land.sum() == synth.N_LAND_HALF_DEG and lat_c.sum() == w.ny
    assert n_land.max() - n_land.min() <= 2 * w.nx   # balanced to within a couple of rows
    for r in range(nranks):
        ys = slice(int(lat_s[r]) - 1, int(lat_s[r]) - 1 + int(lat_c[r]))
        sub = w.window(1, int(lat_s[r]), w.nx, int(lat_c[r]))
        fs = {k: np.ascontiguousarray(v[:, ys, :]) for k, v in f.items()}
        rc, st, a, _ = run(sub, fs, block=THREAD_PER_CELL)
        assert rc == 0
        for n in ("h2osoi_liq", "zwt", "wa", "lai", "plant_mass", "smp", "rootr_col"):
            assert np.array_equal(getattr(st, n)[sub.land], getattr(whole, n)[ys][sub.land]), (r, n)
        assert np.array_equal(a["rnf"], ann["rnf"][ys], equal_nan=True)


def test_four_way_shards_with_the_automatic_shape_reproduce_the_whole(big):
    """What a 4-GPU run does by default: a 16.9k-cell band goes out as one 128-thread block per
    SM of the all-register thread-per-cell build.  Still thread per cell, so bit for bit the
    unsplit run."""
    w, f = big
    _, whole, ann, _ = run(w, f)
    lat_s, lat_c, n_land = partition_lat_bands(w.soil_tex, w.theta_s, 4)
    for r in range(4):
        ys = slice(int(lat_s[r]) - 1, int(lat_s[r]) - 1 + int(lat_c[r]))
        sub = w.window(1, int(lat_s[r]), w.nx, int(lat_c[r]))
        fs = {k: np.ascontiguousarray(v[:, ys, :]) for k, v in f.items()}
        h = make_gpu(sub, mode=MATH_FAST)
        assert h.kernel_variant() == "h9::days_kernel_fast<128,1>", h.kernel_variant()
        h.set_state(init_state(sub.soil_tex, sub.theta_s, synth.ZI_DRIVER))
        assert h.run_days(np.ones(ND, np.int32), fs) == 0
        st = h.get_state()
        h.close()
        for n in ("h2osoi_liq", "zwt", "wa", "lai", "plant_mass", "smp"):
            assert np.array_equal(getattr(st, n)[sub.land], getattr(whole, n)[ys][sub.land]), (r, n)


def test_eight_way_shards_with_the_automatic_kernel_agree_with_the_whole(big):
    """What an 8-GPU run does by default: each 8.4k-cell band steps with two lanes per cell
    (two-sided tridiagonal solve), the unsplit grid with one thread per cell.  Same
    arithmetic up to the order of the solve and of the column sums: after 6 days x 48
    sub-steps the soil water agrees to 1e-6 relative in the median, 99.9 % within 1e-3,
    the water table within 1 mm for 99.9 % of the cells, no fault on either side."""
    w, f = big
    _, whole, ann, _ = run(w, f)
    lat_s, lat_c, n_land = partition_lat_bands(w.soil_tex, w.theta_s, 8)
    rels, dz = [], []
    for r in range(8):
        ys = slice(int(lat_s[r]) - 1, int(lat_s[r]) - 1 + int(lat_c[r]))
        sub = w.window(1, int(lat_s[r]), w.nx, int(lat_c[r]))
        fs = {k: np.ascontiguousarray(v[:, ys, :]) for k, v in f.items()}
        h = make_gpu(sub, mode=MATH_FAST)
        assert "pair" in h.kernel_variant()
        h.set_state(init_state(sub.soil_tex, sub.theta_s, synth.ZI_DRIVER))
        assert h.run_days(np.ones(ND, np.int32), fs) == 0
        st = h.get_state()
        h.close()
        a = st.h2osoi_liq[sub.land].astype(np.float64)
        b = whole.h2osoi_liq[ys][sub.land].astype(np.float64)
        rels.append((np.abs(a - b) / np.maximum(np.abs(b), 1e-3)).ravel())
        dz.append(np.abs(st.zwt[sub.land].astype(np.float64) - whole.zwt[ys][sub.land]))
    rel, dz = np.concatenate(rels), np.concatenate(dz)
    assert np.median(rel) < 1e-6 and np.quantile(rel, 0.999) < 1e-3, (np.median(rel), np.quantile(rel, 0.999))
    assert np.quantile(dz, 0.999) < 1e-3, np.quantile(dz, 0.999)


def test_multi_decade_spin_up_runs_clean():
    """BASELINE.json config 4 in miniature: 30 simulated years from the INIT state on the
    full 0.5 deg land mask, fast mode, one fused launch per year.  No cell may trip one of
    the reference's STOP conditions, and the water table must stay in its legal range.
    (Cells whose water table hovers at the base of the soil column exercise the
    ill-conditioned aquifer-layer formula HYDROLOGY.f90:576-590.)"""
    w = synth.make_world()
    nd = 365
    f = synth.make_forcing(w, nd, seed=9)
    h = make_gpu(w, mode=MATH_FAST, nyr=2)
    h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), with_smp=False)
    p, ds, ps = h.pack_forcing(f, nd)
    for yr in range(30):
        rc = h.run_days_device(np.full(nd, yr % 2 + 1, np.int32), p, ds, ps)
        assert rc == 0, (yr + 1, h.get_fault())
    st = h.get_state()
    land = w.land
    assert (st.zwt[land] >= 0).all() and (st.zwt[land] <= 80).all()
    assert np.isfinite(st.h2osoi_liq[land]).all() and (st.plant_mass[land] > 0).all()
    shallow = (st.zwt[land] <= 2.296).mean()
    assert 0.0 < shallow < 1.0   # both Drainage regimes are populated after the spin-up
    h.close()
