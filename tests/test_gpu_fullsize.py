"""Size-independent properties at BASELINE.json's full size (0.5 deg global land mask,
67,420 cells), where the oracle would take minutes: the reference's in-code invariants,
determinism, and the sharding property behind the multi-GPU layout (cells are
independent, so any latitude-band split reproduces the unsplit run bit for bit)."""
import numpy as np
import pytest

from helpers import assert_state_equal, make_gpu
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


def run(w, f, st=None):
    h = make_gpu(w, mode=MATH_FAST)
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
    w, f = big
    _, whole, ann, _ = run(w, f)
    lat_s, lat_c, n_land = partition_lat_bands(w.soil_tex, w.theta_s, nranks)
    assert n_land.sum() == synth.N_LAND_HALF_DEG and lat_c.sum() == w.ny
    assert n_land.max() - n_land.min() <= 2 * w.nx   # balanced to within a couple of rows
    for r in range(nranks):
        ys = slice(int(lat_s[r]) - 1, int(lat_s[r]) - 1 + int(lat_c[r]))
        sub = w.window(1, int(lat_s[r]), w.nx, int(lat_c[r]))
        fs = {k: np.ascontiguousarray(v[:, ys, :]) for k, v in f.items()}
        rc, st, a, _ = run(sub, fs)
        assert rc == 0
        for n in ("h2osoi_liq", "zwt", "wa", "lai", "plant_mass", "smp", "rootr_col"):
            assert np.array_equal(getattr(st, n)[sub.land], getattr(whole, n)[ys][sub.land]), (r, n)
        assert np.array_equal(a["rnf"], ann["rnf"][ys], equal_nan=True)
