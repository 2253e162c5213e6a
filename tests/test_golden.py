"""Committed fixture tests/golden/h9_golden_v1.npz (made by tests/golden/make_golden.py from
the oracle; the reference itself has no golden vectors): the oracle must keep reproducing it
bit for bit, and the GPU must match it within the stated tolerances."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402
from helpers import STATE_FIELDS, assert_state_close, make_gpu  # noqa: E402
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth  # noqa: E402
from hybrid9_b200.state import H9State  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "h9_golden_v1.npz"))


def test_oracle_reproduces_the_committed_fixture():
    now = make_golden.build()
    assert sorted(now) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(now[k], GOLD[k], equal_nan=True), k


def state_from(prefix):
    return H9State(**{n: np.ascontiguousarray(GOLD[f"{prefix}_{n}"]) for n in
                      ("h2osoi_liq", "zwt", "wa", "lai", "lai_litter", "plant_mass", "plant_foliage_mass",
                       "plant_length", "rdepth", "rootr_col", "nplants", "smp")})


@pytest.mark.gpu
@pytest.mark.parametrize("mode,rtol,atol", [(MATH_EXACT, 2e-3, 0.01), (MATH_FAST, 2e-2, 0.2)])
@pytest.mark.parametrize("tag", ["init", "random"])
def test_gpu_matches_the_committed_fixture(tag, mode, rtol, atol):
    w = synth.make_world(nx=make_golden.NX, ny=make_golden.NY, seed=make_golden.SEED,
                         n_class13=2, n_zero_theta=2)
    f = synth.make_forcing(w, make_golden.NDAYS, seed=make_golden.SEED)
    h = make_gpu(w, nisurf=make_golden.NISURF, nyr=2, mode=mode)
    h.set_state(state_from(f"{tag}_in"))
    rc = h.run_days(GOLD["year_index"], f)
    assert rc == int(GOLD[f"{tag}_rc"]) == 0
    got, ref = h.get_state(), state_from(f"{tag}_out")
    land = w.land
    assert_state_close(got, ref, land, rtol=rtol, atol=atol, fields=("h2osoi_liq", "wa"))
    assert_state_close(got, ref, land, rtol=rtol, atol=2e-3, fields=("zwt", "lai", "lai_litter", "rootr_col"))
    assert_state_close(got, ref, land, rtol=rtol, atol=1e-2, fields=("plant_mass", "plant_foliage_mass"))
    for iy in (1, 2):
        ann = h.get_annual(iy)
        for k in ("npp", "plant_mass", "rnf", "theta_total", "theta"):
            a, b = ann[k][land].astype(np.float64), GOLD[f"{tag}_axy{iy}_{k}"][land].astype(np.float64)
            assert (np.abs(a - b) <= atol * 0.1 + rtol * np.abs(b)).all(), (iy, k, np.abs(a - b).max())
        assert np.all(ann["evap"][land] == 0)
    h.close()
