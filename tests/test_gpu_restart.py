"""Checkpoint / restart (absent in the reference, `notes.txt:12` lists it as
a TODO): h9_get_state -> new context -> h9_set_state continues a run bit for bit when the
restart happens on a year boundary (the annual accumulators restart with every year,
HYBRID9.f90:134-146)."""
import numpy as np
import pytest

from helpers import assert_state_equal, make_gpu
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [MATH_EXACT, MATH_FAST])
def test_restart_on_a_year_boundary_is_bit_exact(mode, tmp_path):
    w = synth.make_world(nx=72, ny=36, seed=9)
    nd = 20
    f = synth.make_forcing(w, 2 * nd, seed=4)
    yi = np.concatenate([np.full(nd, 1, np.int32), np.full(nd, 2, np.int32)])
    a = make_gpu(w, nyr=2, mode=mode)
    a.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), with_smp=False)
    assert a.run_days(yi, f) == 0
    b = make_gpu(w, nyr=2, mode=mode)
    b.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), with_smp=False)
    assert b.run_days(yi[:nd], {k: np.ascontiguousarray(v[:nd]) for k, v in f.items()}) == 0
    ck = b.get_state()
    # the checkpoint goes through the flat restart file of hybrid9_b200/spinup.py
    from hybrid9_b200 import load_restart, save_restart
    save_restart(str(tmp_path / "restart.h9"), ck)
    ck = load_restart(str(tmp_path / "restart.h9"))
    b.close()
    c = make_gpu(w, nyr=2, mode=mode)
    c.set_state(ck, with_smp=True)   # smp is part of the state
    assert c.run_days(yi[nd:], {k: np.ascontiguousarray(v[nd:]) for k, v in f.items()}) == 0
    assert_state_equal(c.get_state(), a.get_state(), w.land)
    for k, v in a.get_annual(2).items():
        assert np.array_equal(v, c.get_annual(2)[k], equal_nan=True), k
    a.close()
    c.close()


def test_spin_up_controller_converges():
    """N3: cycling one forcing year until the land-mean annual soil water drifts by < 0.5 mm per
    cycle; the drift shrinks and the controller stops before its cycle limit."""
    from hybrid9_b200 import spin_up
    w = synth.make_world(nx=72, ny=36, seed=9)
    f = synth.make_forcing(w, 365, seed=4)
    h = make_gpu(w, nyr=1, mode=MATH_FAST)
    h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), with_smp=False)
    hist = spin_up(h, np.ones(365, np.int32), f, w.land, max_cycles=60, tol_mm=0.5)
    drifts = [d for _, _, d in hist[1:]]
    assert len(hist) < 60 and drifts[-1] < 0.5
    assert drifts[-1] < drifts[0]
    h.close()
