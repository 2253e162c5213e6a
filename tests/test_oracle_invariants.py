"""The oracle's own consistency: loop-order independence, thread-count independence,
the reference's in-code invariants, the quirks listed in SURVEY.md appendix A, and
the FP32 rounding-noise floor (float vs double build of the same source)."""
import numpy as np
import pytest

from helpers import assert_state_equal, make_oracle
from hybrid9_b200 import synth


@pytest.fixture(scope="module")
def setup():
    w = synth.make_world(nx=36, ny=18, seed=3)
    nd = 40
    f = synth.make_forcing(w, nd, seed=4)
    return w, f, nd


def run(w, f, nd, yi=None, **kw):
    o = make_oracle(w, nyr=3, **kw)
    o.init_state()
    rc = o.run_days(np.ones(nd, np.int32) if yi is None else yi, f)
    return o, rc


def test_loop_orders_agree_bitwise(setup):
    """cell-outer (HYBRID9.f90:120-295) == time-outer once smp is per cell."""
    w, f, nd = setup
    a, rca = run(w, f, nd, loop_order=0)
    b, rcb = run(w, f, nd, loop_order=1)
    assert rca == 0 and rcb == 0
    assert_state_equal(a.get_state(), b.get_state(), w.land)
    for k, v in a.get_annual(1).items():
        assert np.array_equal(v, b.get_annual(1)[k], equal_nan=True), k


def test_threads_do_not_change_results(setup):
    w, f, nd = setup
    a, _ = run(w, f, nd, loop_order=0, nthreads=1)
    b, _ = run(w, f, nd, loop_order=0, nthreads=4)
    assert_state_equal(a.get_state(), b.get_state(), w.land)


def test_smp_leak_is_a_small_documented_deviation(setup):
    """With the reference's shared smp scratch (SHARED.f90:198) only the first sub-step of
    each cell sees a foreign smp; after 40 days the states agree to rounding level."""
    w, f, nd = setup
    a, _ = run(w, f, nd, loop_order=0, smp_leak=0)
    b, _ = run(w, f, nd, loop_order=0, smp_leak=1)
    sa, sb = a.get_state(), b.get_state()
    land = w.land
    rel = np.abs(sa.h2osoi_liq[land] - sb.h2osoi_liq[land]) / np.abs(sa.h2osoi_liq[land])
    assert rel.max() < 5e-3
    assert np.abs(sa.zwt[land] - sb.zwt[land]).max() < 1e-3


def test_reference_invariants_hold(setup):
    w, f, nd = setup
    o, rc = run(w, f, nd)
    assert rc == 0 and o.get_fault()["n_faulted"] == 0  # |w1-w0| <= 0.1 mm every step, :1244
    st = o.get_state()
    land = w.land
    assert (st.zwt[land] >= 0).all() and (st.zwt[land] <= 80).all()      # :1122-1123
    assert (st.wa[land] <= 5000).all()                                   # :1054
    assert (st.h2osoi_liq[land] >= np.float32(0.01) * (1 - 1e-6)).all()  # watmin repair :1161-1205
    dz = np.array([45, 46, 75, 123, 204, 336, 554, 913], np.float32)
    cap = np.maximum(np.float32(0.01), w.theta_s) * dz
    assert (st.h2osoi_liq[land][:, 1:] <= cap[land][:, 1:] * (1 + 1e-6)).all()  # :1131-1137
    assert (st.lai[land] >= np.float32(0.001)).all()                     # GROW.f90:163
    assert np.all(st.rootr_col[land][:, 8] == 0)                         # G26
    assert np.allclose(st.rootr_col[land].sum(axis=1), 1.0, atol=2e-2)   # 90%-in-rdepth profile


def test_annual_outputs_and_quirks(setup):
    w, f, nd = setup
    yi = np.concatenate([np.full(25, 1, np.int32), np.full(nd - 25, 2, np.int32)])
    o, _ = run(w, f, nd, yi=yi)
    land = w.land
    a1, a2, a3 = o.get_annual(1), o.get_annual(2), o.get_annual(3)
    assert np.all(a1["evap"][land] == 0.0) and np.all(a2["evap"][land] == 0.0)  # G21: axy_evap == 0
    assert np.isnan(a1["npp"][~land]).all() and np.all(a1["theta_total"][~land] == 0)  # INIT.f90:402-414
    assert np.isnan(a3["npp"]).all()          # year never reached: caller's fill survives
    assert np.isfinite(a1["theta"][land]).all() and np.isfinite(a2["rnf"][land]).all()
    # year 1 closed after 25 days: same as a 25-day run
    f25 = {k: np.ascontiguousarray(v[:25]) for k, v in f.items()}
    o25, _ = run(w, f25, 25)
    for k, v in o25.get_annual(1).items():
        assert np.array_equal(v, a1[k], equal_nan=True), k
    # plant_mass mean = sum of end-of-day masses / nt; theta_total = mean column water (G20)
    assert (a1["plant_mass"][land] > 0).all() and (a1["theta_total"][land] > 0).all()


def test_split_calls_equal_one_call(setup):
    """h9_run_days may be called tile by tile: state and accumulators persist."""
    w, f, nd = setup
    a, _ = run(w, f, nd)
    b = make_oracle(w, nyr=3)
    b.init_state()
    for d0 in range(0, nd, 7):
        d1 = min(nd, d0 + 7)
        b.run_days(np.ones(d1 - d0, np.int32), {k: np.ascontiguousarray(v[d0:d1]) for k, v in f.items()})
    assert_state_equal(a.get_state(), b.get_state(), w.land)
    for k, v in a.get_annual(1).items():
        assert np.array_equal(v, b.get_annual(1)[k], equal_nan=True), k


def test_land_mask_edge_cases():
    w = synth.make_world(nx=36, ny=18, seed=7, n_class13=5, n_zero_theta=5)
    o = make_oracle(w)
    idx = o.land_index()
    land = w.land
    assert np.array_equal(idx, np.flatnonzero(land.ravel()))   # y outer, x inner, ascending
    assert (w.soil_tex == 13).sum() == 5 and not land[w.soil_tex == 13].any()
    zero = (w.soil_tex > 0) & (w.soil_tex != 13) & (w.theta_s.sum(axis=2) == 0)
    assert zero.sum() == 5 and not land[zero].any()


def test_empty_and_single_cell_blocks():
    w = synth.make_world(nx=36, ny=18, seed=3)
    # a block of open ocean
    yy, xx = np.nonzero(~w.land)
    e = w.window(int(xx[0]) + 1, int(yy[0]) + 1, 1, 1)
    o = make_oracle(e)
    o.init_state()
    f = synth.make_forcing(e, 2, seed=1, land_only=False)
    assert o.num_land == 0 and o.run_days(np.ones(2, np.int32), f) == 0
    # a single land cell: config[0] of BASELINE.json
    yy, xx = np.nonzero(w.land)
    s = w.window(int(xx[0]) + 1, int(yy[0]) + 1, 1, 1)
    o = make_oracle(s)
    o.init_state()
    f = synth.make_forcing(s, 365, seed=1)
    assert o.num_land == 1 and o.run_days(np.ones(365, np.int32), f) == 0
    a = o.get_annual(1)
    assert np.isfinite(a["npp"]).all() and a["rnf"][0, 0] >= 0


def test_rounding_noise_floor(setup):
    """float build vs double build of the same source: the scale against which the
    GPU tolerances are set (tests/test_gpu_parity.py)."""
    w, f, nd = setup
    a, _ = run(w, f, nd, kind="f32")
    b, _ = run(w, f, nd, kind="f64")
    sa, sb = a.get_state(), b.get_state()
    land = w.land
    rel = np.abs(sa.h2osoi_liq[land] - sb.h2osoi_liq[land]) / np.abs(sb.h2osoi_liq[land])
    assert rel.max() < 2e-3, rel.max()
    assert np.abs(sa.zwt[land] - sb.zwt[land]).max() < 2e-3
