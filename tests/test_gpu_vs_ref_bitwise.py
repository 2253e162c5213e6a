"""GPU exact mode against the reference's own Fortran, BIT FOR BIT.

oracle/_ref/libh9ref_pk.so is the translated reference (oracle/f2cpp.py: HYDROLOGY.f90, GROW.f90,
the loop nest of HYBRID9.f90, statement for statement) compiled with EXP / LOG / real**real mapped
to the same portable double-precision kernels the GPU's exact mode uses (h9::MathExact) instead of
glibc's.  With the three library functions identical on both sides there is no tolerance left:
every state variable, every annual mean, every diagnostic and the fault record must be equal.
(`smp` is per cell on both sides; the reference's shared-scratch leak is covered on the CPU.)
The library is built where /root/reference exists and travels to the GPU box prebuilt."""
import numpy as np
import pytest

import ref_py
from helpers import STATE_FIELDS, assert_state_equal, day_slice, make_gpu
from hybrid9_b200 import MATH_EXACT, calendar, synth
from hybrid9_b200.state import init_state

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_py.available("pk"), reason="oracle/_ref/libh9ref_pk.so not present")]
ALL = STATE_FIELDS + ("nplants",)


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def world():
    return synth.make_world(nx=144, ny=72, seed=5, n_class13=5, n_zero_theta=5)


@pytest.fixture(scope="module")
def forcing(world):
    return synth.make_forcing(world, 12, seed=3)


def states(world):
    st = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    return {"init": st, "random": synth.randomize_state(world, st, seed=11)}


def test_land_index(world):
    h = make_gpu(world, mode=MATH_EXACT)
    r = ref_py.make_ref(world, kind="pk")
    assert h.num_land == r.num_land and same(h.land_index(), r.land_index())
    h.close()


def test_single_hydrology_calls(world, forcing):
    """12 consecutive CALL HYDROLOGY from states with the water table in every layer."""
    st0 = states(world)["random"]
    h = make_gpu(world, mode=MATH_EXACT)
    r = ref_py.make_ref(world, kind="pk")
    h.set_state(st0)
    r.set_state(st0)
    land = world.land
    stopped = np.zeros(land.shape, bool)
    for step in range(12):
        f = day_slice(forcing, step)
        go, ro = h.hydrology_step(f), r.hydrology_step(f)
        stopped |= np.abs(ro["w_imbalance"]) > 0.1  # the reference defines nothing after its STOP
        ok = land & ~stopped
        for k in ("theta", "qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc", "w_imbalance", "jwt"):
            assert same(go[k][ok], ro[k][ok]), (step, k)
        assert_state_equal(h.get_state(), r.get_state(), ok, fields=("h2osoi_liq", "smp", "zwt", "wa"),
                           what=f"step {step}: ")
    assert stopped.sum() < 0.2 * land.sum()
    h.close()


def test_grow(world, forcing):
    st0 = states(world)["random"]
    h = make_gpu(world, mode=MATH_EXACT)
    r = ref_py.make_ref(world, kind="pk")
    h.set_state(st0)
    r.set_state(st0)
    land = world.land
    for d in range(4):
        tas = np.ascontiguousarray(forcing["tas"][d] + np.float32(5.0 * d - 8.0))
        go, ro = h.grow_day(tas), r.grow_day(tas)
        for k in ("npp", "w_i", "fT"):
            assert same(go[k][land], ro[k][land]), (d, k)
        assert_state_equal(h.get_state(), r.get_state(), land, fields=ALL, what=f"day {d}: ")
    h.close()


@pytest.mark.parametrize("tag", ["init", "random"])
def test_fused_day_kernel_twelve_days(world, forcing, tag):
    """K3 through h9_run_days against HYBRID9.f90:126-292 cell by cell: 12 days x 48 sub-steps +
    GROW over two year slots; state and the annual means of both years."""
    st0 = states(world)[tag]
    yi = np.concatenate([np.full(7, 1, np.int32), np.full(5, 2, np.int32)])
    h = make_gpu(world, nyr=2, mode=MATH_EXACT)
    r = ref_py.make_ref(world, nyr=2, kind="pk")
    h.set_state(st0)
    r.set_state(st0)
    rc_r = r.run_days(yi, forcing)
    assert rc_r == 0, "the reference STOPs from this state; pick another seed"
    assert h.run_days(yi, forcing) == 0
    land = world.land
    assert_state_equal(h.get_state(), r.get_state(), land, fields=ALL)
    for iy in (1, 2):
        ga, ra = h.get_annual(iy), r.get_annual(iy)
        for k in ga:
            assert same(ga[k], ra[k]), (iy, k)  # whole grid: land values and the NaN / 0 fills
    h.close()


def test_anchor_year(world):
    """BASELINE.json configs[0]: a single-cell-scale block, calendar year 1901 (365 d x 48)."""
    w = synth.make_world(nx=4, ny=2, seed=3)
    nd = calendar.time_boy(1902) - calendar.time_boy(1901)
    f = synth.make_forcing(w, nd, seed=3)
    st0 = init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER)
    h = make_gpu(w, mode=MATH_EXACT)
    r = ref_py.make_ref(w, kind="pk")
    h.set_state(st0, with_smp=False)
    r.set_state(st0, with_smp=False)
    yi = np.ones(nd, np.int32)
    assert r.run_days(yi, f) == 0 and h.run_days(yi, f) == 0
    assert_state_equal(h.get_state(), r.get_state(), w.land, fields=ALL)
    ga, ra = h.get_annual(1), r.get_annual(1)
    for k in ga:
        assert same(ga[k], ra[k]), k
    h.close()


def test_stop_record(world, forcing):
    """NISURF = 1 trips the reference's water-balance STOP; same cell, day, sub-step, imbalance."""
    st0 = states(world)["random"]
    f = {k: np.ascontiguousarray(v[:1]) for k, v in forcing.items()}
    h = make_gpu(world, nisurf=1, mode=MATH_EXACT)
    r = ref_py.make_ref(world, nisurf=1, kind="pk")
    h.set_state(st0)
    r.set_state(st0)
    assert r.run_days(np.ones(1, np.int32), f) == 8
    rc = h.run_days(np.ones(1, np.int32), f)
    gf, rf = h.get_fault(), r.get_fault()
    assert rc & 8
    assert (gf.code, gf.x, gf.y, gf.day, gf.substep) == (rf["code"], rf["x"], rf["y"], rf["day"], rf["substep"])
    assert np.float32(gf.imbalance) == np.float32(rf["imbalance"])
    h.close()
