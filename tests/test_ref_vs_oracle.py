"""THE PIN: the hand-written oracle (oracle/h9_oracle.cpp) against the reference's own Fortran,
translated statement for statement by oracle/f2cpp.py and compiled here (oracle/_ref/libh9ref.so;
HYDROLOGY.f90 and GROW.f90 whole, INIT.f90 / HYBRID9.f90 by line range).

Both sides are FP32, strict IEEE evaluation order, the same glibc powf / expf / logf, so the
comparison is BIT-EXACT everywhere: layer geometry, calendar, initial state, land mask and
iteration order, single HYDROLOGY calls from states that put the water table in every layer
(all 17 compared outputs and locals), GROW, multi-day runs with per-cell `smp` and with the
reference's shared `smp` scratch (the leak), annual means, a whole 3652-day decade on the
reference's calendar through the verbatim loop nest, and the STOP conditions.

Runs where the library exists: it is built in this container from /root/reference and travels
to the GPU box as a prebuilt file; without either the module is skipped."""
import os

import numpy as np
import pytest

import oracle_py
import ref_py
from helpers import STATE_FIELDS, assert_state_equal, day_slice, make_oracle
from hybrid9_b200 import calendar, synth
from hybrid9_b200.state import geometry, init_state, land_mask

pytestmark = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref/libh9ref.so not built "
                                "(needs /root/reference) and no prebuilt copy present")

ALL_FIELDS = STATE_FIELDS + ("nplants",)


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def world():
    return synth.make_world(nx=72, ny=36, seed=9, n_class13=5, n_zero_theta=5)


@pytest.fixture(scope="module")
def forcing(world):
    return synth.make_forcing(world, 12, seed=4)


def test_translated_units_are_the_reference_files():
    """The manifest the translator wrote names every unit of the hot path."""
    path = os.path.join(ref_py.REF_DIR, "manifest.txt")
    if not os.path.exists(path):
        pytest.skip("manifest is only written where the translator ran")
    text = open(path).read()
    for unit in ("HYDROLOGY.f90: SUBROUTINE HYDROLOGY", "GROW.f90: SUBROUTINE GROW",
                 "HYBRID9.f90:120-295 -> decade_loop()", "INIT.f90:711-811 -> init_state()",
                 "INIT.f90:844-859 -> init_time_boy()", "SHARED.f90: whole module",
                 "CONTROL.f90: whole module"):
        assert unit in text, unit


def test_geometry_calendar_and_initial_state(world):
    r = ref_py.make_ref(world)
    o = make_oracle(world)
    for a, b in zip(r.geometry(), o.geometry()):
        assert same(np.float32(a), np.float32(b))
    dt, dz, zc = geometry(synth.ZI_DRIVER, 48)
    assert same(r.geometry()[0], dz) and same(r.geometry()[1], zc) and r.geometry()[2] == dt
    lib = oracle_py.load()
    for year in range(1860, 2301):
        assert r.time_boy(year) == lib.h9o_time_boy(year) == calendar.time_boy(year), year
    r.init_state()
    o.init_state()
    land = world.land
    assert_state_equal(r.get_state(), o.get_state(), land, fields=ALL_FIELDS)
    assert_state_equal(r.get_state(), init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER),
                       land, fields=ALL_FIELDS)


def test_land_mask_and_iteration_order(world, forcing):
    """The verbatim loop nest (HYBRID9.f90:120-295) writes axy_* exactly for the cells the
    oracle, the host library and hybrid9_b200.state.land_mask call land."""
    r = ref_py.make_ref(world, per_cell_smp=False)
    r.init_state()
    f1 = {k: v[:1] for k, v in forcing.items()}
    assert r.run_days(np.ones(1, np.int32), f1) == 0
    processed = ~np.isnan(r.get_annual(1)["npp"])
    assert same(processed, world.land) and same(processed, land_mask(world.soil_tex, world.theta_s))
    o = make_oracle(world)
    assert same(r.land_index(), o.land_index())
    assert same(np.flatnonzero(processed.ravel()).astype(np.int32), o.land_index())
    assert (~processed).sum() >= 10  # class-13 and zero-porosity cells are in the block


@pytest.mark.parametrize("leak", [0, 1])
def test_multi_day_run_from_init_state(world, forcing, leak):
    """leak=1: the reference as it is -- ONE smp scratch vector for all cells (SHARED.f90:198),
    verbatim loop nest.  leak=0: smp per cell (what the GPU does)."""
    nd = 12
    yi = np.concatenate([np.full(7, 1, np.int32), np.full(nd - 7, 2, np.int32)])
    r = ref_py.make_ref(world, nyr=2, per_cell_smp=not leak)
    o = make_oracle(world, nyr=2, loop_order=0, smp_leak=leak)
    r.init_state()
    o.init_state()
    assert r.run_days(yi, forcing) == 0 and o.run_days(yi, forcing) == 0
    land = world.land
    fields = tuple(n for n in ALL_FIELDS if not (leak and n == "smp"))
    assert_state_equal(r.get_state(), o.get_state(), land, fields=fields)
    for iy in (1, 2):
        ar, ao = r.get_annual(iy), o.get_annual(iy)
        for k in ar:
            assert same(ar[k][land], ao[k][land]), (iy, k)
            if k != "theta_total":
                assert np.isnan(ar[k][~land]).all(), k  # INIT.f90:402-413
        assert (ar["theta_total"][~land] == 0).all()    # INIT.f90:414
        assert (ar["evap"][land] == 0).all()            # evap_sum is never accumulated


def test_annual_forcing_means(world, forcing):
    """HYBRID9.f90:235-241,277-283: FP32 running sums in day order, divided by nt."""
    r = ref_py.make_ref(world, per_cell_smp=False)
    r.init_state()
    nd = 12
    assert r.run_days(np.ones(nd, np.int32), forcing) == 0
    got = r.get_annual_forcing(1)
    land = world.land
    for k in ref_py.FORCING:
        s = np.zeros(land.shape, np.float32)
        for d in range(nd):
            s = (s + forcing[k][d].astype(np.float32)).astype(np.float32)
        assert same(got[k][land], (s / np.float32(nd))[land]), k


def test_single_hydrology_calls_with_the_water_table_in_every_layer(world, forcing):
    """24 consecutive CALL HYDROLOGY from randomised states: every output and 16 locals."""
    st0 = synth.randomize_state(world, init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), seed=21)
    r = ref_py.make_ref(world)
    o = make_oracle(world)
    r.set_state(st0)
    o.set_state(st0)
    land = world.land
    yy, xx = np.nonzero(land)
    zi_m = synth.ZI_DRIVER.astype(np.float32) / np.float32(1000.0)
    jwt0 = np.array([next((i - 1 for i in range(1, 9) if z <= zi_m[i]), 8) for z in st0.zwt[land]])
    assert set(jwt0.tolist()) == set(range(9))  # HYDROLOGY.f90:499-508: every case is entered
    seen_jwt = set()
    faulted = np.zeros(land.shape, bool)
    for step in range(24):
        f = day_slice(forcing, step % 12)
        ro, oo = r.hydrology_step(f), o.hydrology_step(f)
        faulted |= (np.abs(oo["w_imbalance"]) > 0.1)  # nothing is defined after the reference's STOP
        ok = land & ~faulted
        for k in ("theta", "qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc", "w_imbalance", "jwt"):
            assert same(ro[k][ok], oo[k][ok]), (step, k)
        assert_state_equal(r.get_state(), o.get_state(), ok, what=f"step {step}: ")
        seen_jwt |= set(np.unique(oo["jwt"][ok]).tolist())
        if step % 6 == 0:  # the oracle exposes its locals one cell at a time
            for j in range(0, yy.size, 7):
                y, x = int(yy[j]), int(xx[j])
                if faulted[y, x]:
                    continue
                d = o.step_diag(x + 1, y + 1)
                for name in ("qflx_surf", "rsub_top", "qflx_rsub_sat", "qflx_infl", "qcharge", "fsat",
                             "beta", "rsc", "w0", "w1"):
                    assert np.float32(getattr(d, name)) == ro["diag"][name][y, x], (step, x, y, name)
    assert len(seen_jwt) >= 5, seen_jwt
    assert faulted.sum() < 0.2 * land.sum()


def test_grow(world, forcing):
    st0 = synth.randomize_state(world, init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), seed=22)
    r = ref_py.make_ref(world)
    o = make_oracle(world)
    r.set_state(st0)
    o.set_state(st0)
    land = world.land
    for d in range(6):
        tas = forcing["tas"][d] + np.float32(4.0 * d - 8.0)  # both branches of fT about 18 C
        ro, oo = r.grow_day(tas), o.grow_day(tas)
        for k in ("npp", "w_i", "fT"):
            assert same(ro[k][land], oo[k][land]), (d, k)
        assert_state_equal(r.get_state(), o.get_state(), land, what=f"day {d}: ")


def test_randomised_states_ten_days(world, forcing):
    st0 = synth.randomize_state(world, init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), seed=23)
    r = ref_py.make_ref(world)
    o = make_oracle(world, loop_order=0)
    # NISURF = 48 from these states trips no STOP (checked); faults are covered below
    r.set_state(st0)
    o.set_state(st0)
    yi = np.ones(10, np.int32)
    f = {k: v[:10] for k, v in forcing.items()}
    rr, orc = r.run_days(yi, f), o.run_days(yi, f)
    assert rr == orc == 0
    assert_state_equal(r.get_state(), o.get_state(), world.land, fields=ALL_FIELDS)
    ar, ao = r.get_annual(1), o.get_annual(1)
    for k in ar:
        assert same(ar[k][world.land], ao[k][world.land]), k


def test_stop_conditions(world, forcing):
    """NISURF = 1 from randomised states trips |w1-w0| > 0.1 mm (HYDROLOGY.f90:1244-1273): the
    translated reference stops at the same cell, day and sub-step, with the same imbalance, as
    the oracle's first-fault record."""
    st0 = synth.randomize_state(world, init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), seed=11)
    r = ref_py.make_ref(world, nisurf=1)
    o = make_oracle(world, nisurf=1, loop_order=0)
    r.set_state(st0)
    o.set_state(st0)
    f = {k: v[:1] for k, v in forcing.items()}
    rr, orc = r.run_days(np.ones(1, np.int32), f), o.run_days(np.ones(1, np.int32), f)
    assert rr == 8 and (orc & 8), (rr, orc)
    rf, of = r.get_fault(), o.get_fault()
    assert (rf["code"], rf["x"], rf["y"], rf["day"], rf["substep"]) == \
        (of["code"], of["x"], of["y"], of["day"], of["substep"])
    assert np.float32(rf["imbalance"]) == np.float32(of["imbalance"])
    assert rf["line"] == 1273


@pytest.mark.parametrize("idec,idec_start", [(1, 1), (12, 12)])
def test_whole_decade_on_the_reference_calendar(idec, idec_start):
    """PROGRAM H9's decade exactly as written: HYBRID9.f90:103-113 (syr, eyr) and the loop
    nest :120-295 over time_BOY, 3652 days x 48 sub-steps (731 days for decade 12)."""
    w = synth.make_world(nx=16, ny=8, seed=31)
    land = w.land
    assert 8 <= land.sum() <= 40
    nd = calendar.decade_days(idec)
    assert nd == (3652 if idec == 1 else 731)
    f = synth.make_forcing(w, nd, seed=6)
    yi = calendar.year_index_of_days(idec, idec_start)
    nyr = int(yi[-1])
    r = ref_py.make_ref(w, nyr=nyr, per_cell_smp=False)
    o = make_oracle(w, nyr=nyr, loop_order=0, smp_leak=1)
    r.init_state()
    o.init_state()
    assert r.run_decade(idec_start, idec, f) == 0 and o.run_days(yi, f) == 0
    assert_state_equal(r.get_state(), o.get_state(), land,
                       fields=tuple(n for n in ALL_FIELDS if n != "smp"))
    for iy in range(1, nyr + 1):
        ar, ao = r.get_annual(iy), o.get_annual(iy)
        for k in ar:
            assert same(ar[k][land], ao[k][land]), (iy, k)


def test_bounds_checked_build_runs_clean(world, forcing):
    """Every subscript checked (the -fcheck=bounds of this build): no out-of-range access on
    either the INIT state or randomised states, and the same bits as the unchecked build."""
    if not ref_py.available("chk"):
        pytest.skip("bounds-checked build not present")
    st0 = synth.randomize_state(world, init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), seed=24)
    f = {k: v[:2] for k, v in forcing.items()}
    out = []
    for kind in ("chk", "ref"):
        r = ref_py.make_ref(world, kind=kind)
        r.set_state(st0)
        assert r.run_days(np.ones(2, np.int32), f) == 0
        out.append(r.get_state())
    assert_state_equal(out[0], out[1], world.land, fields=ALL_FIELDS)


@pytest.mark.parametrize("tag", ["init", "random"])
def test_kernel_source_equals_the_translated_reference_under_the_same_math(world, forcing, tag):
    """The GPU's exact-mode physics (hybrid9_b200/csrc/h9_physics.h compiled for the host,
    tests/twin) against the translated reference built with the SAME portable pow/exp/log
    (libh9ref_pk.so): bit for bit.  The GPU reproduces that host build bit for bit
    (tests/test_gpu_parity.py) and is compared with libh9ref_pk.so directly in
    tests/test_gpu_vs_ref_bitwise.py."""
    if not ref_py.available("pk"):
        pytest.skip("libh9ref_pk.so not present")
    st0 = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    if tag == "random":
        st0 = synth.randomize_state(world, st0, seed=23)
    nd = 8
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing.items()}
    r = ref_py.make_ref(world, kind="pk")
    r.set_state(st0)
    assert r.run_days(np.ones(nd, np.int32), f) == 0
    tw, ex = oracle_py.twin_run(world, st0, f, 48, synth.ZI_DRIVER, math="exact")
    assert not ex["fault"].any()
    assert_state_equal(r.get_state(), tw, world.land)
