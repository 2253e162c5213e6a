"""Host-side logic that stays outside the GPU path: calendar, INIT's initial state,
the synthetic world, array conventions."""
import numpy as np

import oracle_py
from helpers import assert_state_equal, make_oracle
from hybrid9_b200 import calendar, synth
from hybrid9_b200.state import H9State, geometry, init_state, land_mask


def test_time_boy_matches_the_oracle_and_known_answers():
    tb = oracle_py.load("f32").h9o_time_boy
    for y in (1860, 1861, 1900, 1901, 1904, 1905, 2000, 2001, 2012, 2013, 2100, 2300):
        assert calendar.time_boy(y) == tb(y)
    assert calendar.time_boy(1901) == 14976
    assert calendar.decade_days(1) == 3652 and calendar.decade_days(12) == 731
    assert sum(calendar.decade_days(d) for d in range(1, 13)) == 40908


def test_year_index_of_days():
    yi = calendar.year_index_of_days(1)
    assert yi.size == 3652 and yi[0] == 1 and yi[-1] == 10
    assert np.array_equal(np.bincount(yi)[1:], [365, 365, 365, 366, 365, 365, 365, 366, 365, 365])
    yi2 = calendar.year_index_of_days(2, idec_start=1)
    assert yi2[0] == 11 and yi2[-1] == 20
    assert calendar.year_index_of_days(3, idec_start=3)[0] == 1   # iY is relative to iDEC_start


def test_init_state_bitexact_vs_oracle():
    w = synth.make_world(nx=72, ny=36, seed=9)
    o = make_oracle(w)
    o.init_state()
    assert_state_equal(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), o.get_state(), w.land)


def test_geometry_matches_init():
    dt, dz, zc = geometry(synth.ZI_DRIVER, 48)
    assert dt == 1800.0 and dz[9] == 2704.0 and zc[9] == 3648.0


def test_synthetic_world_has_the_documented_shape():
    w = synth.make_world()   # 0.5 deg
    assert (w.nx, w.ny) == (720, 360)
    assert int(w.land.sum()) == synth.N_LAND_HALF_DEG == 67420
    assert (w.soil_tex == 13).sum() == 200
    assert w.lat[0] == 89.75 and w.lon[0] == -179.75          # INIT.f90:142-145
    land = w.land
    assert w.theta_s[land].min() >= 0.30 and w.theta_s[land].max() <= 0.55
    assert w.psi_s[land].max() < 0 and w.bsw[land].min() > 2.8 and w.bsw[land].max() < 12.6
    assert not land[(w.lat > 84) | (w.lat < -56)].any()
    assert np.array_equal(land, land_mask(w.soil_tex, w.theta_s))
    # regional window of BASELINE.json config 2: ~2-3k land cells
    r = w.window(*synth.REGIONAL_WINDOW)
    assert int(r.land.sum()) == 2500
    # determinism
    w2 = synth.make_world()
    assert np.array_equal(w.soil_tex, w2.soil_tex) and np.array_equal(w.bsw, w2.bsw)


def test_forcing_shape_and_ranges():
    w = synth.make_world(nx=72, ny=36, seed=9)
    f = synth.make_forcing(w, 20, seed=1)
    land = w.land
    for k in ("tas", "rlds", "rsds", "huss", "ps", "pr", "rhs"):
        assert f[k].shape == (20, 36, 72) and f[k].dtype == np.float32 and f[k].flags["C_CONTIGUOUS"]
        assert np.all(f[k][:, ~land] == 0)
    assert f["tas"][:, land].min() >= 220 and f["tas"][:, land].max() <= 320
    assert f["rsds"][:, land].min() >= 0 and f["pr"][:, land].min() >= 0
    assert 0.2 < (f["pr"][:, land] > 0).mean() < 0.4
    assert f["rhs"][:, land].min() >= 5 and f["rhs"][:, land].max() <= 100


def test_array_conventions_are_the_fortran_ones():
    """numpy C-order (lat_c, lon_c, 8) is byte-identical to Fortran (8, lon_c, lat_c)."""
    st = H9State.zeros(3, 5)
    st.h2osoi_liq[2, 4, 7] = 1.0
    flat = st.h2osoi_liq.ravel()
    assert flat[(2 * 5 + 4) * 8 + 7] == 1.0     # ((y-1)*lon_c + (x-1))*8 + (I-1)
    cw = synth.compact_world(synth.make_world(nx=72, ny=36, seed=9), 10, start=3)
    assert (cw.nx, cw.ny) == (10, 1) and cw.land.all()
