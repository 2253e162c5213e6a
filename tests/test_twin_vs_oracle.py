"""The kernel source (hybrid9_b200/csrc/h9_physics.h), compiled for the host with
the C library's powf/expf/logf, must reproduce the independently written oracle
BIT FOR BIT: two restatements of HYDROLOGY.f90 / GROW.f90 with different code
structure (hoisted per-day terms, predicated loops, register arrays vs the
Fortran's statement order).  Any difference is a logic bug in one of them.
The same source with the portable exact-mode kernels is then compared to the
oracle within a stated tolerance; the GPU's exact mode reproduces THAT build
bit for bit (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import oracle_py
from helpers import assert_state_equal, day_slice, make_oracle
from hybrid9_b200 import synth


@pytest.fixture(scope="module")
def world():
    return synth.make_world(nx=144, ny=72, seed=5)


def test_init_state_30_days_bitexact(world):
    w = world
    f = synth.make_forcing(w, 30, seed=3)
    o = make_oracle(w)
    o.init_state()
    st0 = o.get_state()
    assert o.run_days(np.ones(30, np.int32), f) == 0
    tw, ex = oracle_py.twin_run(w, st0, f, 48, synth.ZI_DRIVER, math="libm")
    assert not ex["fault"].any()
    assert_state_equal(tw, o.get_state(), w.land)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_states_every_branch_bitexact(world, seed):
    """Randomised states put the water table in every layer (jwt = 0..8), layers from
    dry to over-saturated, aquifer at its cap, canopy from bare to closed."""
    w = world
    f = synth.make_forcing(w, 3, seed=seed)
    o = make_oracle(w)
    o.init_state()
    st0 = synth.randomize_state(w, o.get_state(), seed=seed)
    # one sub-step with full diagnostics
    o.set_state(st0)
    out = o.hydrology_step(day_slice(f, 0))
    tw1, ex1 = oracle_py.twin_run(w, st0, {k: v[:1] for k, v in f.items()}, 1, synth.ZI_DRIVER,
                                  do_grow=False, math="libm")
    land = w.land
    # nisurf=1 changes dt, so redo the oracle with nisurf=1 for the one-step comparison
    o1 = make_oracle(w, nisurf=1)
    o1.set_state(st0)
    out = o1.hydrology_step(day_slice(f, 0))
    assert_state_equal(tw1, o1.get_state(), land, what="1 step: ")
    assert np.array_equal(ex1["jwt"], out["jwt"][land])
    assert np.array_equal(ex1["theta"], out["theta"][land])
    for k in ("qflx_tran_veg_col", "qflx_evap_grnd", "w_imbalance"):
        a, b = ex1[k], out[k][land]
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all(), k
    # the inputs visit every water-table position: jwt = 0..8 (HYDROLOGY.f90:499-508)
    jwt_in = (st0.zwt[land][:, None] > synth.ZI_DRIVER[None, 1:9] / np.float32(1000.0)).sum(axis=1)
    assert (np.bincount(jwt_in, minlength=9) > 0).all()
    # three days with GROW
    o.set_state(st0)
    rc = o.run_days(np.ones(3, np.int32), f)
    tw, ex = oracle_py.twin_run(w, st0, f, 48, synth.ZI_DRIVER, math="libm")
    ok = land.copy()
    ok[land] = ex["fault"] == 0  # a faulted cell is past the reference's STOP: undefined
    assert int((ex["fault"] != 0).sum()) == o.get_fault()["n_faulted"]
    assert rc == int(np.bitwise_or.reduce(ex["fault"]))
    assert_state_equal(tw, o.get_state(), ok, what="3 days: ")


def test_grow_day_bitexact(world):
    w = world
    o = make_oracle(w)
    o.init_state()
    st0 = synth.randomize_state(w, o.get_state(), seed=21)
    f = synth.make_forcing(w, 4, seed=5)
    o.set_state(st0)
    o.run_days(np.ones(4, np.int32), f)
    tw, ex = oracle_py.twin_run(w, st0, f, 48, synth.ZI_DRIVER, math="libm")
    # last day's GROW diagnostics from a fresh GROW call on the same pre-state are covered by
    # the state equality; here check the diagnostics path of the oracle entry itself
    o2 = make_oracle(w)
    o2.set_state(st0)
    g = o2.grow_day(f["tas"][0])
    land = w.land
    assert np.isfinite(g["npp"][land]).all()
    assert ((g["w_i"][land] >= 0) & (g["w_i"][land] <= 1.0 + 1e-6)).all()
    assert (g["fT"][land] <= 1).all()
    assert_state_equal(tw, o.get_state(), land)


def test_portable_kernels_are_correctly_rounded():
    lib = oracle_py.load_twin()
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.uniform(0.01, 1.0, 3000), rng.uniform(1.0, 400.0, 3000)]).astype(np.float32)
    b = rng.uniform(-30, 30, a.size).astype(np.float32)
    got = np.array([lib.h9t_pow(float(x), float(y)) for x, y in zip(a, b)], np.float32)
    with np.errstate(over="ignore"):
        ref = np.power(a.astype(np.float64), b.astype(np.float64)).astype(np.float32)
    ok = np.isfinite(ref) & (ref > 1e-37)
    assert np.array_equal(got[ok], ref[ok])
    x = rng.uniform(-80, 80, 3000).astype(np.float32)
    assert np.array_equal(np.array([lib.h9t_exp(float(v)) for v in x], np.float32),
                          np.exp(x.astype(np.float64)).astype(np.float32))
    y = np.exp(rng.uniform(-80, 80, 3000)).astype(np.float32)
    assert np.array_equal(np.array([lib.h9t_log(float(v)) for v in y], np.float32),
                          np.log(y.astype(np.float64)).astype(np.float32))
    assert lib.h9t_pow(0.0, 2.0) == 0.0 and lib.h9t_pow(2.0, 0.0) == 1.0
    assert np.isinf(lib.h9t_pow(0.0, -2.0)) and np.isnan(lib.h9t_pow(-1.0, 0.5))
    assert lib.h9t_pow(0.01, 2e8) == 0.0 and np.isinf(lib.h9t_pow(0.01, -2e8))  # bsw = 1e8 (G28)


def test_portable_kernels_corner_regions():
    """The places where the table-driven kernels change regime: bases next to 1 (centre c = 1
    of the log table), mantissas either side of the sqrt(2) split in every binade, sub-normal and
    largest floats, saturating exponents, and the special values of the wrappers."""
    lib = oracle_py.load_twin()
    rng = np.random.default_rng(7)
    n = 4000

    def check(a, b):
        got = np.array([lib.h9t_pow(float(x), float(y)) for x, y in zip(a, b)], np.float32)
        with np.errstate(all="ignore"):
            ref = np.power(a.astype(np.float64), b.astype(np.float64)).astype(np.float32)
        assert np.array_equal(got, ref), np.nonzero(got != ref)[0][:5]

    check((1 + rng.uniform(-1e-3, 1e-3, n)).astype(np.float32), rng.uniform(-3000, 3000, n).astype(np.float32))
    split = (np.float32(1.41421) + rng.uniform(-2e-5, 2e-5, n)).astype(np.float32)
    check(split * np.exp2(rng.integers(-20, 20, n)).astype(np.float32), rng.uniform(-3, 3, n).astype(np.float32))
    check(np.exp(rng.uniform(-103, 88, n)).astype(np.float32), rng.uniform(-3, 3, n).astype(np.float32))
    check(np.array([1e-45, 1.1754944e-38, 3.4028235e38, 1.0, 0.5, 2.0], np.float32),
          np.array([0.5, 0.5, 0.5, 7.0, 200.0, -200.0], np.float32))
    inf, nan = float("inf"), float("nan")
    assert np.isnan(lib.h9t_exp(nan)) and np.isinf(lib.h9t_exp(inf)) and lib.h9t_exp(-inf) == 0.0
    assert np.isinf(lib.h9t_pow(2.0, inf)) and lib.h9t_pow(0.5, inf) == 0.0 and lib.h9t_pow(2.0, -inf) == 0.0
    assert lib.h9t_pow(1.0, inf) == 1.0 and lib.h9t_pow(inf, -1.0) == 0.0 and np.isinf(lib.h9t_pow(inf, 1.0))
    assert np.isnan(lib.h9t_pow(nan, 1.0)) and np.isnan(lib.h9t_pow(2.0, nan)) and lib.h9t_pow(nan, 0.0) == 1.0
    assert np.isinf(lib.h9t_log(0.0)) and np.isnan(lib.h9t_log(-1.0)) and np.isinf(lib.h9t_log(inf))
    x = np.array([-104.0, -103.9, -87.4, 88.7, 88.8, 0.0, -0.0, 1e-30], np.float32)
    with np.errstate(over="ignore"):
        ref = np.exp(x.astype(np.float64)).astype(np.float32)
    assert np.array_equal(np.array([lib.h9t_exp(float(v)) for v in x], np.float32), ref)


def test_exact_mode_build_vs_oracle_tolerance(world):
    """Exact-mode arithmetic differs from the oracle only in pow/exp/log results that
    differ by one float ulp on rare inputs (glibc's powf is not always correctly
    rounded).  Stated tolerance for 10 days from randomised states: 99.9 % of the soil
    water values within 5e-5 relative, every value within 2e-3 relative + 0.01 mm."""
    w = world
    f = synth.make_forcing(w, 10, seed=3)
    o = make_oracle(w)
    o.init_state()
    st0 = synth.randomize_state(w, o.get_state(), seed=11)
    o.set_state(st0)
    o.run_days(np.ones(10, np.int32), f)
    ref = o.get_state()
    tw, ex = oracle_py.twin_run(w, st0, f, 48, synth.ZI_DRIVER, math="exact")
    land = w.land
    a, b = tw.h2osoi_liq[land].astype(np.float64), ref.h2osoi_liq[land].astype(np.float64)
    rel = np.abs(a - b) / np.abs(b)
    assert np.quantile(rel, 0.999) < 5e-5, np.quantile(rel, 0.999)
    assert (np.abs(a - b) <= 0.01 + 2e-3 * np.abs(b)).all(), rel.max()
    assert np.abs(tw.zwt[land] - ref.zwt[land]).max() < 2e-3


@pytest.mark.parametrize("nisurf", [1, 24, 172])
def test_other_substep_counts_bitexact(world, nisurf):
    """dt = 86400/NISURF (INIT.f90:214; the reference's notes ran 1, 48 and 172)."""
    w = world
    f = synth.make_forcing(w, 2, seed=31)
    o = make_oracle(w, nisurf=nisurf)
    o.init_state()
    st0 = o.get_state()
    rc = o.run_days(np.ones(2, np.int32), f)
    tw, ex = oracle_py.twin_run(w, st0, f, nisurf, synth.ZI_DRIVER, math="libm")
    assert rc == int(np.bitwise_or.reduce(ex["fault"]))
    ok = w.land.copy()
    ok[w.land] = ex["fault"] == 0
    assert_state_equal(tw, o.get_state(), ok)
