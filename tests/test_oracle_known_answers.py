"""Pins the oracle against the hand-derived known answers of the reference's INIT
formulas.  The reference ships no tests or golden vectors
(parity is UNPINNED by the reference); these values were derived by hand from
INIT.f90:214,252-257,707-811,844-859 and EXECUTE/driver.txt:2,17-26."""
import numpy as np
import pytest

import oracle_py
from helpers import make_oracle
from hybrid9_b200 import synth


@pytest.fixture(scope="module")
def small():
    w = synth.make_world(nx=36, ny=18, seed=2)
    o = make_oracle(w)
    o.init_state()
    return w, o, o.get_state()


def test_geometry(small):
    _, o, _ = small
    dz, zc, dt = o.geometry()
    assert dt == 1800.0  # 86400/48, INIT.f90:214
    assert list(dz[1:]) == [45, 46, 75, 123, 204, 336, 554, 913, 2704]
    assert list(zc[1:]) == [22.5, 68, 128.5, 227.5, 391, 661, 1106, 1839.5, 3648]


def test_initial_state(small):
    w, o, st = small
    land = w.land
    assert land.sum() > 0 and o.num_land == land.sum()
    dz = np.array([45, 46, 75, 123, 204, 336, 554, 913], np.float32)
    expect = (np.float32(0.4) * w.theta_s * dz * np.float32(1000.0) / np.float32(1000.0))[land]
    assert np.array_equal(st.h2osoi_liq[land], expect)           # INIT.f90:730
    assert np.all(st.zwt[land] == np.float32(7.296))             # (2296+5000)/1000, INIT.f90:739
    assert np.all(st.wa[land] == 4000.0)                         # INIT.f90:744
    assert np.all(st.lai_litter[land] == np.float32(0.001))      # INIT.f90:748
    assert np.all(st.nplants[land] == 1)
    assert np.allclose(st.plant_length[land], 50.3058, rtol=2e-6)  # (400/3.142e-3)**(1/3)
    assert np.allclose(st.rdepth[land], 15.0917, rtol=5e-6)  # 0.3 * plant_length
    assert np.allclose(st.lai[land], 0.0435 * 0.023, rtol=1e-6)  # 1.0005e-3
    r = st.rootr_col[land][0]
    assert abs(r[0] - 0.998957) < 2e-6 and abs(r[1] - 1.04195e-3) < 2e-7
    assert abs(r[2] - 9.337e-7) < 1.2e-7  # one float ulp of (1 - decay**z) near 1
    assert np.all(r[3:] == 0.0)
    assert np.all(st.h2osoi_liq[~land] == 0) and np.all(st.wa[~land] == 0)


def test_initial_water_table_is_below_the_column(small):
    """zwt0 = 7.296 m => jwt0 = 8, zc(9) = 4567.75, dz(9) = 5456.5 (HYDROLOGY.f90:499-508,645-650)."""
    w, o, _ = small
    o2 = make_oracle(w)
    o2.init_state()
    f = synth.make_forcing(w, 1, seed=1)
    out = o2.hydrology_step({k: v[0] for k, v in f.items()})
    yy, xx = np.nonzero(w.land)
    d = o2.step_diag(int(xx[0]) + 1, int(yy[0]) + 1)
    assert d.jwt_soilwater == 8
    assert out["fault"] == 0
    assert np.all(out["jwt"][w.land] == 8)


def test_calendar():
    tb = oracle_py.load("f32").h9o_time_boy
    assert tb(1860) == 1
    assert tb(1901) == 14976                    # SURVEY.md 8c; notes: PGF 1901 starts Time=14975 (0-based)
    assert tb(1911) - tb(1901) == 3652          # days in 1901-1910
    assert tb(2013) - tb(1901) == 40908         # days in 1901-2012
    assert tb(1905) - tb(1904) == 366 and tb(1901) - tb(1900) == 365  # 1900 is not a leap year
    assert tb(2001) - tb(2000) == 366
