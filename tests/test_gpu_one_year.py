"""One-year trajectory (365 days x 48 sub-steps = 17,520 HYDROLOGY calls + 365 GROW calls per
cell) of 1,500 land cells of the 0.5 deg world: GPU vs oracle, with the tolerances SURVEY.md
section 8c states for a year (theta rel 1e-3 / abs 1e-4, zwt abs 1e-3 m, plant mass and LAI
rel 1e-3, annual runoff rel 1e-3), applied to the bulk of the cells; the tail of ill-conditioned
cells is bounded by 10x the FP32 rounding-noise floor of the model (float vs double oracle)."""
import os

import numpy as np
import pytest

from helpers import make_gpu, make_oracle
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu
N, ND = 1500, 365


@pytest.fixture(scope="module")
def year():
    w = synth.make_world()
    f = synth.make_forcing(w, ND, seed=9)
    cw = synth.compact_world(w, N, start=30000)
    cf = synth.compact_forcing(w, f, N, start=30000)
    out = {}
    for kind in ("f32", "f64"):
        o = make_oracle(cw, kind=kind, loop_order=0, nthreads=os.cpu_count() or 1)
        o.init_state()
        assert o.run_days(np.ones(ND, np.int32), cf) == 0
        out[kind] = (o.get_state(), o.get_annual(1))
    return cw, cf, out


def rel(a, b, floor):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


@pytest.mark.parametrize("mode,bulk", [(MATH_EXACT, 0.999), (MATH_FAST, 0.99)])
def test_one_year_against_the_oracle(year, mode, bulk):
    cw, cf, out = year
    ref, ann = out["f32"]
    ref64, ann64 = out["f64"]
    h = make_gpu(cw, mode=mode)
    h.set_state(init_state(cw.soil_tex, cw.theta_s, synth.ZI_DRIVER), with_smp=False)
    assert h.run_days(np.ones(ND, np.int32), cf) == 0
    got, gann = h.get_state(), h.get_annual(1)
    dz = np.array([45, 46, 75, 123, 204, 336, 554, 913], np.float64)
    th_g, th_r, th_64 = got.h2osoi_liq / dz, ref.h2osoi_liq / dz, ref64.h2osoi_liq / dz
    checks = [  # (name, gpu, oracle, double oracle, rel tol, abs floor)
        ("theta", th_g, th_r, th_64, 1e-3, 1e-1),
        ("plant_mass", got.plant_mass, ref.plant_mass, ref64.plant_mass, 1e-3, 1.0),
        ("lai", got.lai, ref.lai, ref64.lai, 1e-3, 1e-2),
        ("axy_rnf", gann["rnf"], ann["rnf"], ann64["rnf"], 1e-3, 1e-5),
        ("axy_theta", gann["theta"], ann["theta"], ann64["theta"], 1e-3, 1e-1),
        ("axy_npp", gann["npp"], ann["npp"], ann64["npp"], 1e-3, 1.0),
    ]
    for name, g, r, r64, tol, floor in checks:
        e = rel(g, r, floor)
        noise = rel(r, r64, floor)
        assert np.quantile(e, bulk) < tol, (name, float(np.quantile(e, bulk)))
        assert e.max() < 10 * max(noise.max(), tol), (name, float(e.max()), float(noise.max()))
    dzw = np.abs(got.zwt.astype(np.float64) - ref.zwt)
    assert np.quantile(dzw, bulk) < 1e-3
    assert dzw.max() < 10 * max(np.abs(ref.zwt.astype(np.float64) - ref64.zwt).max(), 1e-3)
    h.close()
