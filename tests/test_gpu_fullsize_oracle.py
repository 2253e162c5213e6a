"""Parity with the oracle at BASELINE.json's full sizes (`-m gpu`).

configs[2]: the whole 0.5 deg land mask (67,420 cells) x 10 days x 48 sub-steps, exact and
fast mode (thread-per-cell kernel: the one that steps this size), against the oracle on all
host threads (about 3e7 cell-steps: seconds of CPU).
configs[4]: a band of the 0.25 deg mask (1440 x 720 grid, 269,680 land cells; rows around
the equator and a mid-latitude band, ~40k cells) the same way.
configs[3]: the observable of a spin-up is its equilibrium: 30 simulated years from
randomised states (water table in every layer, both Drainage regimes) on a 2.7k-cell block,
fast mode (both stepping kernels) against the GPU's exact mode, which is the reference's
arithmetic bit for bit (tests/test_gpu_vs_ref_bitwise.py): per-cell annual means of year 30
and the end state within stated tolerance, and no bias of the land mean.
Tolerances are the ones of tests/test_gpu_parity.py (tied to the FP32 noise floor there)."""
import os

import numpy as np
import pytest

from helpers import THREAD_PER_CELL, TWO_LANES, assert_state_close, make_gpu, make_oracle
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu
ND = 10
NTHREADS = os.cpu_count() or 1


def rel(a, b, floor=1e-3):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def oracle_run(w, st, yi, f):
    o = make_oracle(w, nthreads=NTHREADS)
    o.set_state(st)
    assert o.run_days(yi, f) == 0
    out = (o.get_state(), o.get_annual(1))
    o.close()
    return out


def check_vs_oracle(w, f, mode, tag):
    st = init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER)
    yi = np.ones(ND, np.int32)
    ref, oa = oracle_run(w, st, yi, f)
    h = make_gpu(w, mode=mode, block=THREAD_PER_CELL if mode == MATH_FAST else 0)
    assert np.array_equal(h.land_index(), np.flatnonzero(w.land.ravel()))  # A0: bit-exact
    h.set_state(st)
    assert h.run_days(yi, f) == 0
    got, ga = h.get_state(), h.get_annual(1)
    h.close()
    land = w.land
    r = rel(got.h2osoi_liq[land], ref.h2osoi_liq[land])
    if mode == MATH_EXACT:
        # differs from the oracle only where glibc's powf is not correctly rounded
        assert np.quantile(r, 0.999) < 1e-4, (tag, np.quantile(r, 0.999))
        assert_state_close(got, ref, land, rtol=5e-3, atol=0.02, fields=("h2osoi_liq", "wa"))
        assert_state_close(got, ref, land, rtol=0, atol=2e-3, fields=("zwt",))
    else:
        assert np.median(r) < 2e-5 and np.quantile(r, 0.999) < 5e-3, (tag, np.median(r), np.quantile(r, 0.999))
        assert_state_close(got, ref, land, rtol=5e-2, atol=0.5, fields=("h2osoi_liq", "wa"))
        assert np.abs(got.zwt[land] - ref.zwt[land]).max() < 2e-2
    # GROW switches its foliage-loss law at w_i = 0.6 (GROW.f90:136-138, G24): a cell that sits on
    # the switch flips with the last bit of w_i, in any arithmetic, and its LAI jumps by 10 % of the
    # foliage.  So: 99.9 % of the cells tightly, every cell within one such jump.
    for n in ("lai", "plant_mass", "rootr_col"):
        a, b = getattr(got, n)[land].astype(np.float64), getattr(ref, n)[land].astype(np.float64)
        e = np.abs(a - b) / np.maximum(np.abs(b), 1e-4)
        assert np.quantile(e, 0.999) < (2e-2 if mode == MATH_FAST else 2e-3), (tag, n, np.quantile(e, 0.999))
        assert e.max() < 0.2, (tag, n, e.max())
    for k, at in (("npp", 1e-3), ("plant_mass", 1e-4), ("rnf", 1e-7), ("theta_total", 0.05), ("theta", 1e-5)):
        rt = 2e-3 if mode == MATH_EXACT else 1e-2
        d = np.abs(ga[k][land].astype(np.float64) - oa[k][land])
        e = d - (at + rt * np.abs(oa[k][land]))
        if mode == MATH_EXACT:
            assert (e <= 0).all(), (tag, k, e.max())
        else:  # 99.9 % of the cells within the gate, every cell within one GROW switch (see above)
            assert (e <= 0).mean() > 0.999, (tag, k, (e <= 0).mean())
            assert (d <= at + 0.2 * np.abs(oa[k][land])).all(), (tag, k, d.max())
    assert np.isnan(ga["npp"][~land]).all() and np.all(ga["theta_total"][~land] == 0)


@pytest.mark.parametrize("mode", [MATH_EXACT, MATH_FAST], ids=["exact", "fast"])
def test_half_degree_globe_10_days_vs_oracle(mode):
    w = synth.make_world()
    assert int(w.land.sum()) == synth.N_LAND_HALF_DEG
    check_vs_oracle(w, synth.make_forcing(w, ND, seed=9), mode, "0.5deg")


@pytest.mark.parametrize("mode", [MATH_EXACT, MATH_FAST], ids=["exact", "fast"])
def test_quarter_degree_band_10_days_vs_oracle(mode):
    w = synth.make_world(nx=1440, ny=720, n_land=269680, seed=9)
    assert int(w.land.sum()) == 269680
    n = 0
    for lat_s, rows in ((330, 40), (150, 40)):  # around the equator; northern mid-latitudes
        sub = w.window(1, lat_s, w.nx, rows)
        n += int(sub.land.sum())
        check_vs_oracle(sub, synth.make_forcing(sub, ND, seed=9), mode, f"0.25deg rows {lat_s}+{rows}")
    assert n > 30000


@pytest.fixture(scope="module")
def spinup():
    """30 years x 365 days, the same forcing year cycled (a spin-up), three GPU runs."""
    w = synth.make_world(nx=144, ny=72, seed=5)
    f = synth.make_forcing(w, 365, seed=3)
    st = synth.randomize_state(w, init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), seed=11)
    jwt_in = (st.zwt[w.land][:, None] > synth.ZI_DRIVER[None, 1:9] / np.float32(1000.0)).sum(axis=1)
    assert (np.bincount(jwt_in, minlength=9) > 0).all()
    out = {}
    for name, mode, block in (("exact", MATH_EXACT, 0), ("thread", MATH_FAST, THREAD_PER_CELL),
                              ("pair", MATH_FAST, TWO_LANES)):
        h = make_gpu(w, mode=mode, nyr=2, block=block)
        h.set_state(st)
        p, ds, ps = h.pack_forcing(f, 365)
        faulted = 0
        for yr in range(30):
            rc = h.run_days_device(np.full(365, yr % 2 + 1, np.int32), p, ds, ps)
            faulted |= rc
        out[name] = (h.get_state(), h.get_annual(2), faulted, h.get_fault())
        h.close()
    return w, out


@pytest.mark.parametrize("kernel", ["thread", "pair"])
def test_thirty_year_equilibrium_fast_vs_exact(spinup, kernel):
    """Year 30 of a spin-up from states with the water table in every layer; no run may trip
    one of the reference's STOP conditions."""
    w, out = spinup
    se, ae, fe, ffe = out["exact"]
    sf, af, ff, fff = out[kernel]
    land = w.land
    assert fe == 0 and ff == 0, (ffe, fff)
    ok = land
    shallow = (se.zwt[land] <= 2.296).mean()
    assert 0.02 < shallow < 0.98, shallow   # both Drainage regimes populated at equilibrium
    # per-cell: annual means of year 30.  Gates = the FP32 rounding-noise floor of the model on this
    # very case (oracle float vs double after 30 years, profiles/r02/equilibrium_report_30_years.json:
    # p99 / p99.9 of rnf 3.8e-2 / 0.22, theta 3.1e-3 / 3.1e-2, plant_mass 6e-4 / 3.7e-3,
    # theta_total 3.5e-3 / 1.9e-2, npp 8.8e-2 / 0.63); measured fast vs exact: rnf 1.8e-3 / 1.7e-2,
    # theta 3.5e-4 / 1.3e-2, plant_mass 1.5e-4 / 1.3e-3, theta_total 4.9e-4 / 9.7e-3, npp 6.9e-3 / 0.29
    for k, q99, rt in (("rnf", 2e-2, 0.2), ("theta", 3e-3, 3e-2), ("plant_mass", 6e-4, 4e-3),
                       ("theta_total", 3.5e-3, 2e-2), ("npp", 5e-2, 0.6)):
        a, b = af[k][ok].astype(np.float64), ae[k][ok].astype(np.float64)
        floor = {"rnf": 1e-6, "npp": 1e-3}.get(k, 1e-3)
        r = np.abs(a - b) / np.maximum(np.abs(b), floor)
        assert np.quantile(r, 0.99) < q99, (k, "p99", np.quantile(r, 0.99))
        assert np.quantile(r, 0.999) < rt, (k, "p99.9", np.quantile(r, 0.999))
        # no bias of the land mean
        # no bias of the land mean (noise floor 7e-5 .. 1.7e-4; measured <= 3.6e-4 for npp)
        bias = abs(a.mean() - b.mean()) / max(abs(b.mean()), 1e-12)
        assert bias < 1e-3, (k, "land-mean bias", bias)
    dz = np.abs(sf.zwt[ok].astype(np.float64) - se.zwt[ok])
    assert np.quantile(dz, 0.99) < 2e-2 and abs(sf.zwt[ok].mean() - se.zwt[ok].mean()) < 2e-3, \
        (np.quantile(dz, 0.99), sf.zwt[ok].mean() - se.zwt[ok].mean())
    r = rel(sf.h2osoi_liq[ok], se.h2osoi_liq[ok])
    assert np.median(r) < 1e-4 and np.quantile(r, 0.99) < 2e-2, (np.median(r), np.quantile(r, 0.99))
