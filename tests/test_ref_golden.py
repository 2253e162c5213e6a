"""tests/golden/h9_ref_golden_v1.npz holds outputs of the REFERENCE'S OWN Fortran (translated and
compiled by oracle/f2cpp.py + g++; generator tests/golden/make_ref_golden.py, which needs
/root/reference).  The vectors are committed so that they travel:

* CPU: where /root/reference is present the translated reference must regenerate them bit for
  bit; everywhere, the hand-written oracle must reproduce them bit for bit;
* GPU (`-m gpu`): the CUDA path through the C ABI must match them -- indexing exactly, values
  within the tolerances written here (exact mode: glibc's powf vs the portable correctly
  rounded kernels; fast mode: MUFU-based pow, bounded against the FP32 noise floor elsewhere).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_golden as mg  # noqa: E402
import ref_py  # noqa: E402
from helpers import assert_state_close, make_gpu, make_oracle  # noqa: E402
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth  # noqa: E402
from hybrid9_b200.state import H9State  # noqa: E402

GOLD = np.load(mg.PATH)
ANNUAL = ("npp", "plant_mass", "rnf", "evap", "theta_total", "theta")


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def state_from(prefix):
    return H9State(**{n: np.ascontiguousarray(GOLD[f"{prefix}_{n}"]) for n in mg.FIELDS})


@pytest.mark.skipif(not (os.path.isdir(ref_py.REFERENCE_SRC) and ref_py.available()),
                    reason="needs /root/reference to translate and compile the reference")
def test_translated_reference_regenerates_the_fixture():
    now = mg.build()
    assert sorted(now) == sorted(GOLD.files)
    for k in GOLD.files:
        assert same(now[k], GOLD[k]), k


@pytest.mark.parametrize("tag", ["init", "random", "leak"])
def test_oracle_reproduces_the_reference_vectors(tag):
    w = mg.small_world()
    f = synth.make_forcing(w, mg.NDAYS, seed=mg.SEED)
    land = w.land
    o = make_oracle(w, nisurf=mg.NISURF, nyr=2, loop_order=0, smp_leak=int(tag == "leak"))
    o.set_state(state_from(f"{tag}_in"))
    assert o.run_days(mg.year_index(), f) == int(GOLD[f"{tag}_rc"]) == 0
    got = o.get_state()
    for n in mg.FIELDS:
        if tag == "leak" and n == "smp":
            continue  # the module's one scratch vector is not per-cell state
        assert same(getattr(got, n)[land], GOLD[f"{tag}_out_{n}"][land]), n
    for iy in (1, 2):
        ann = o.get_annual(iy)
        for k in ANNUAL:
            assert same(ann[k][land], GOLD[f"{tag}_axy{iy}_{k}"][land]), (iy, k)


def test_oracle_reproduces_the_reference_single_calls():
    w = mg.small_world()
    f = synth.make_forcing(w, mg.NDAYS, seed=mg.SEED)
    f0 = {k: np.ascontiguousarray(v[0]) for k, v in f.items()}
    land = w.land
    o = make_oracle(w, nisurf=mg.NISURF)
    o.set_state(state_from("step_in"))
    out = o.hydrology_step(f0)
    assert out["fault"] == int(GOLD["step_fault"])
    for k in ("theta", "qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc", "w_imbalance", "jwt"):
        assert same(out[k][land], GOLD[f"step_{k}"][land]), k
    got = o.get_state()
    for n in mg.FIELDS:
        assert same(getattr(got, n)[land], GOLD[f"step_out_{n}"][land]), n
    g = o.grow_day(f0["tas"])
    for k in ("npp", "w_i", "fT"):
        assert same(g[k][land], GOLD[f"grow_{k}"][land]), k
    got = o.get_state()
    for n in mg.FIELDS:
        assert same(getattr(got, n)[land], GOLD[f"grow_out_{n}"][land]), n


def year_case():
    from hybrid9_b200 import calendar
    from hybrid9_b200.state import init_state
    wy = mg.year_world()
    nd = mg.year_days()
    fy = synth.make_forcing(wy, calendar.decade_days(1), seed=mg.YSEED)
    return wy, nd, {k: np.ascontiguousarray(v[:nd]) for k, v in fy.items()}, \
        init_state(wy.soil_tex, wy.theta_s, synth.ZI_DRIVER)


def test_oracle_reproduces_the_reference_year():
    """BASELINE.json configs[0]: a single-cell-scale block, one calendar year."""
    wy, nd, f, st0 = year_case()
    assert nd == 365
    land = wy.land
    assert land.sum() >= 1
    o = make_oracle(wy, nisurf=mg.NISURF, nyr=1, loop_order=0)
    o.set_state(st0, with_smp=False)
    assert o.run_days(np.ones(nd, np.int32), f) == int(GOLD["year_rc"]) == 0
    got = o.get_state()
    for n in mg.FIELDS:
        assert same(getattr(got, n)[land], GOLD[f"year_out_{n}"][land]), n
    ann = o.get_annual(1)
    for k in ANNUAL:
        assert same(ann[k][land], GOLD[f"year_axy1_{k}"][land]), k


# ---- GPU -----------------------------------------------------------------------------------------

TOL = {MATH_EXACT: dict(rtol=2e-3, atol=0.01), MATH_FAST: dict(rtol=2e-2, atol=0.2)}


def check_gpu_state(got, prefix, land, rtol, atol):
    ref = state_from(prefix)
    assert_state_close(got, ref, land, rtol=rtol, atol=atol, fields=("h2osoi_liq", "wa"))
    assert_state_close(got, ref, land, rtol=rtol, atol=2e-3, fields=("zwt", "lai", "lai_litter", "rootr_col"))
    assert_state_close(got, ref, land, rtol=rtol, atol=1e-2, fields=("plant_mass", "plant_foliage_mass"))
    assert np.array_equal(got.nplants[land], ref.nplants[land])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [MATH_EXACT, MATH_FAST])
@pytest.mark.parametrize("tag", ["init", "random"])
def test_gpu_matches_the_reference_vectors(tag, mode):
    w = mg.small_world()
    f = synth.make_forcing(w, mg.NDAYS, seed=mg.SEED)
    land = w.land
    rtol, atol = TOL[mode]["rtol"], TOL[mode]["atol"]
    h = make_gpu(w, nisurf=mg.NISURF, nyr=2, mode=mode)
    assert np.array_equal(h.land_index(), np.flatnonzero(~np.isnan(GOLD["init_axy1_npp"]).ravel()))
    h.set_state(state_from(f"{tag}_in"))
    assert h.run_days(mg.year_index(), f) == int(GOLD[f"{tag}_rc"]) == 0
    check_gpu_state(h.get_state(), f"{tag}_out", land, rtol, atol)
    for iy in (1, 2):
        ann = h.get_annual(iy)
        for k in ("npp", "plant_mass", "rnf", "theta_total", "theta"):
            a, b = ann[k][land].astype(np.float64), GOLD[f"{tag}_axy{iy}_{k}"][land].astype(np.float64)
            assert (np.abs(a - b) <= atol * 0.1 + rtol * np.abs(b)).all(), (iy, k, np.abs(a - b).max())
            assert np.isnan(ann[k][~land]).all() or k == "theta_total"
        assert np.all(ann["evap"][land] == 0)
    h.close()


@pytest.mark.gpu
def test_gpu_single_calls_match_the_reference_vectors():
    """One CALL HYDROLOGY and one CALL GROW from the branch-coverage state (exact mode): jwt
    exact, values to 1e-5 relative (single-step tolerance of SURVEY.md section 8c)."""
    w = mg.small_world()
    f = synth.make_forcing(w, mg.NDAYS, seed=mg.SEED)
    f0 = {k: np.ascontiguousarray(v[0]) for k, v in f.items()}
    land = w.land
    h = make_gpu(w, nisurf=mg.NISURF, mode=MATH_EXACT)
    h.set_state(state_from("step_in"))
    out = h.hydrology_step(f0)
    ok = land & (np.abs(GOLD["step_w_imbalance"]) <= 0.05)
    assert np.array_equal(out["jwt"][ok], GOLD["step_jwt"][ok])
    for k, floor in (("theta", 1e-6), ("qflx_tran_veg_col", 1e-9), ("qflx_evap_grnd", 1e-9), ("rnf_inc", 1e-7)):
        a, b = out[k][ok].astype(np.float64), GOLD[f"step_{k}"][ok].astype(np.float64)
        assert (np.abs(a - b) <= 1e-4 * np.abs(b) + floor).all(), (k, np.abs(a - b).max())
    got = h.get_state()
    ref = state_from("step_out")
    assert_state_close(got, ref, ok, rtol=1e-5, atol=1e-4, fields=("h2osoi_liq", "wa"))
    assert_state_close(got, ref, ok, rtol=1e-5, atol=1e-6, fields=("zwt",))
    h.set_state(ref)
    g = h.grow_day(f0["tas"])
    for k in ("npp", "w_i", "fT"):
        assert np.allclose(g[k][land], GOLD[f"grow_{k}"][land], rtol=1e-5, atol=1e-6), k
    assert_state_close(h.get_state(), state_from("grow_out"), land, rtol=1e-5, atol=1e-6,
                       fields=("lai", "lai_litter", "plant_mass", "plant_foliage_mass", "plant_length",
                               "rdepth", "rootr_col"))
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [MATH_EXACT, MATH_FAST])
def test_gpu_matches_the_reference_year(mode):
    """configs[0] on the GPU: one year of the anchor block against the reference's vectors with
    SURVEY.md section 8c's one-year tolerances (theta rel 1e-3, zwt 1e-3 m, plant mass / LAI /
    annual runoff rel 1e-3; fast mode 5x)."""
    wy, nd, f, st0 = year_case()
    land = wy.land
    k = 1.0 if mode == MATH_EXACT else 5.0
    h = make_gpu(wy, nisurf=mg.NISURF, nyr=1, mode=mode)
    h.set_state(st0, with_smp=False)
    assert h.run_days(np.ones(nd, np.int32), f) == 0
    got, ref = h.get_state(), state_from("year_out")
    dz = np.array([45, 46, 75, 123, 204, 336, 554, 913], np.float64)
    th_g, th_r = got.h2osoi_liq[land] / dz, ref.h2osoi_liq[land] / dz
    assert (np.abs(th_g - th_r) <= k * (1e-3 * np.abs(th_r) + 1e-4)).all(), np.abs(th_g - th_r).max()
    assert (np.abs(got.zwt[land].astype(np.float64) - ref.zwt[land]) <= k * 1e-3).all()
    for n in ("plant_mass", "lai"):
        a, b = getattr(got, n)[land].astype(np.float64), getattr(ref, n)[land].astype(np.float64)
        assert (np.abs(a - b) <= k * 1e-3 * np.abs(b) + 1e-6).all(), n
    ann = h.get_annual(1)
    for n, floor in (("rnf", 1e-9), ("npp", 1e-3), ("theta", 1e-4), ("plant_mass", 1e-3)):
        a, b = ann[n][land].astype(np.float64), GOLD[f"year_axy1_{n}"][land].astype(np.float64)
        assert (np.abs(a - b) <= k * 1e-3 * np.abs(b) + floor).all(), (n, np.abs(a - b).max())
    h.close()
