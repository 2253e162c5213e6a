"""C-ABI behaviour around the hot path: argument and call-order errors, other NISURF values
(the reference's notes ran 1, 48 and 172 sub-steps per day), mode switching, several contexts
at once, shards with no land at all, re-configuration."""
import ctypes as C

import numpy as np
import pytest

import oracle_py
from helpers import assert_state_close, assert_state_equal, make_gpu, make_oracle
from hybrid9_b200 import H9, MATH_EXACT, MATH_FAST, host, synth
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    return synth.make_world(nx=72, ny=36, seed=9)


def test_call_order_and_argument_errors(world):
    lib = host.load_library()
    h = H9(0)
    z = np.zeros(4, np.float32)
    p = z.ctypes.data_as(host.c_f)
    # before h9_configure / h9_set_soil
    assert lib.h9_set_soil(h.h, None, p, p, p, p, p) == -3          # H9_ERR_STATE
    assert lib.h9_run_days(h.h, 1, None, p, p, p, p, p, p, p) == -3
    assert lib.h9_configure(h.h, 0, 1, 48, p, 1) == -1               # H9_ERR_ARG
    assert lib.h9_configure(h.h, 2, 2, 0, p, 1) == -1
    assert b"bad argument" in lib.h9_last_error(h.h)
    h.configure(world.nx, world.ny, 48, synth.ZI_DRIVER, nyr=1)
    assert lib.h9_set_soil(h.h, None, p, p, p, p, p) == -1           # null pointer
    assert lib.h9_get_annual(h.h, 1, None, None, None, None, None, None) == -3
    h.set_soil(world.soil_tex, world.theta_s, world.hksat, world.bsw, world.psi_s, world.fmax)
    assert lib.h9_get_annual(h.h, 2, None, None, None, None, None, None) == -1   # iyr > nyr
    assert lib.h9_set_math(h.h, 7) == -1
    with pytest.raises(host.H9Error):
        h.set_soil(world.soil_tex[:-1], world.theta_s, world.hksat, world.bsw, world.psi_s, world.fmax)
    h.close()
    h.close()  # idempotent on the Python side


@pytest.mark.parametrize("nisurf", [1, 24, 172])
def test_other_substep_counts(world, nisurf):
    """dt = 86400/NISURF (INIT.f90:214).  Exact mode == host twin bit for bit; fast mode close."""
    nd = 3
    f = synth.make_forcing(world, nd, seed=5)
    st = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    tw, ex = oracle_py.twin_run(world, st, f, nisurf, synth.ZI_DRIVER, math="exact")
    h = make_gpu(world, nisurf=nisurf, mode=MATH_EXACT)
    h.set_state(st, with_smp=False)
    rc = h.run_days(np.ones(nd, np.int32), f)
    assert rc == int(np.bitwise_or.reduce(ex["fault"]))
    ok = world.land.copy()
    ok[world.land] = ex["fault"] == 0
    assert_state_equal(h.get_state(), tw, ok)
    h.close()
    if nisurf > 1:   # NISURF = 1 is a 24 h step: the model itself is outside its stable range
        o = make_oracle(world, nisurf=nisurf)
        o.set_state(st, with_smp=False)
        assert o.run_days(np.ones(nd, np.int32), f) == 0
        g = make_gpu(world, nisurf=nisurf, mode=MATH_FAST)
        g.set_state(st, with_smp=False)
        assert g.run_days(np.ones(nd, np.int32), f) == 0
        assert_state_close(g.get_state(), o.get_state(), world.land, rtol=5e-3, atol=0.05,
                           fields=("h2osoi_liq", "wa", "zwt", "lai", "plant_mass"))
        g.close()


def test_math_mode_can_be_switched_between_runs(world):
    nd = 4
    f = synth.make_forcing(world, nd, seed=6)
    st = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    half = {k: np.ascontiguousarray(v[:2]) for k, v in f.items()}
    rest = {k: np.ascontiguousarray(v[2:]) for k, v in f.items()}
    a = make_gpu(world, mode=MATH_EXACT)
    a.set_state(st, with_smp=False)
    a.run_days(np.ones(2, np.int32), half)
    mid = a.get_state()
    a.set_math(MATH_FAST)
    a.run_days(np.ones(2, np.int32), rest)
    b = make_gpu(world, mode=MATH_FAST)
    b.set_state(mid)
    b.run_days(np.ones(2, np.int32), rest)
    assert_state_equal(a.get_state(), b.get_state(), world.land)
    a.close()
    b.close()


def test_contexts_are_independent(world):
    """Two contexts interleaved == each one alone (one ctx per MPI rank in the reference's layout)."""
    nd = 3
    f = synth.make_forcing(world, nd, seed=7)
    st = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    w2 = world.window(1, 10, world.nx, 12)
    f2 = {k: np.ascontiguousarray(v[:, 9:21, :]) for k, v in f.items()}
    st2 = init_state(w2.soil_tex, w2.theta_s, synth.ZI_DRIVER)
    a, b = make_gpu(world, mode=MATH_FAST), make_gpu(w2, mode=MATH_FAST)
    a.set_state(st, with_smp=False)
    b.set_state(st2, with_smp=False)
    for d in range(nd):
        a.run_days(np.ones(1, np.int32), {k: np.ascontiguousarray(v[d:d + 1]) for k, v in f.items()})
        b.run_days(np.ones(1, np.int32), {k: np.ascontiguousarray(v[d:d + 1]) for k, v in f2.items()})
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa.h2osoi_liq[9:21][w2.land], sb.h2osoi_liq[w2.land])
    assert np.array_equal(sa.zwt[9:21][w2.land], sb.zwt[w2.land])
    a.close()
    b.close()


def test_reconfigure_resets_the_context(world):
    h = make_gpu(world, mode=MATH_FAST)
    h.set_state(init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), with_smp=False)
    f = synth.make_forcing(world, 2, seed=8)
    h.run_days(np.ones(2, np.int32), f)
    small = world.window(1, 10, 36, 12)
    h.configure(small.nx, small.ny, 48, synth.ZI_DRIVER, nyr=1)
    assert h.num_land == -1                      # soil must be set again
    h.set_soil(small.soil_tex, small.theta_s, small.hksat, small.bsw, small.psi_s, small.fmax)
    assert h.num_land == int(small.land.sum())
    h.set_state(init_state(small.soil_tex, small.theta_s, synth.ZI_DRIVER), with_smp=False)
    fs = {k: np.ascontiguousarray(v[:, 9:21, :36]) for k, v in f.items()}
    assert h.run_days(np.ones(2, np.int32), fs) == 0
    assert np.isfinite(h.get_annual(1)["theta"][small.land]).all()
    h.close()


def test_latitude_band_without_land_is_harmless(world):
    """A polar band of an 8-way split holds no land: every entry point must accept nc == 0."""
    polar = world.window(1, 1, world.nx, 1)
    assert not polar.land.any()
    h = make_gpu(polar, mode=MATH_FAST)
    h.set_state(init_state(polar.soil_tex, polar.theta_s, synth.ZI_DRIVER), with_smp=False)
    f = synth.make_forcing(polar, 3, seed=1, land_only=False)
    assert h.run_days(np.ones(3, np.int32), f) == 0
    p, ds, ps = h.pack_forcing(f, 3)
    assert h.run_days_device(np.ones(3, np.int32), p, ds, ps) == 0
    ptr, stride, b = h.annual_device(1)
    h.synchronize()
    assert h.get_fault().any == 0 and h.counters()["launches"] >= 0
    out = h.hydrology_step({k: np.ascontiguousarray(v[0]) for k, v in f.items()})
    assert out["fault"] == 0
    h.close()
