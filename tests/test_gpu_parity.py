"""GPU parity tests (run on the B200 box, `-m gpu`): everything goes through the
C ABI of libh9gpu.so.

Parity chain (the reference ships no golden vectors and cannot be compiled here):
  oracle (oracle/h9_oracle.cpp, follows the Fortran statement by statement)
    == host twin of the kernel source with libm        bit-exact  (test_twin_vs_oracle.py, CPU)
  host twin with the portable exact-mode kernels
    == GPU, H9_MATH_EXACT                              bit-exact  (here)
  GPU, H9_MATH_EXACT vs oracle                         stated tolerance (pow/exp/log differ by
                                                       <= 1 float ulp on rare inputs)
  GPU, H9_MATH_FAST vs oracle                          stated, looser tolerance, compared with
                                                       the FP32 rounding-noise floor (f32 vs f64
                                                       build of the oracle)
Cells past a reference STOP condition (fault word != 0) are excluded from value
comparisons: the reference defines nothing after its STOP.
"""
import json
import os

import numpy as np
import pytest

import oracle_py
from helpers import (FAST_KERNEL_IDS, FAST_KERNELS, THREAD_128REG, THREAD_PER_CELL, TWO_LANES,
                     assert_state_close, assert_state_equal, day_slice, make_gpu, make_oracle)
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATS = {}


def record(name, **kw):
    STATS[name] = {k: (float(v) if np.isscalar(v) else v) for k, v in kw.items()}
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_stats.json"), "w") as f:
            json.dump(STATS, f, indent=1)


def relerr(a, b, floor=1e-3):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


@pytest.fixture(scope="module")
def world():
    return synth.make_world(nx=144, ny=72, seed=5)


@pytest.fixture(scope="module")
def forcing30(world):
    return synth.make_forcing(world, 30, seed=3)


def states(world):
    st_init = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    return st_init, synth.randomize_state(world, st_init, seed=11)


# ---- indexing: bit-exact -------------------------------------------------------


def test_land_mask_and_compaction_order(world):
    h = make_gpu(world)
    o = make_oracle(world)
    assert h.num_land == o.num_land == int(world.land.sum())
    assert np.array_equal(h.land_index(), o.land_index())
    assert np.array_equal(h.land_index(), np.flatnonzero(world.land.ravel()))
    h.close()


def test_state_roundtrip_bitexact(world):
    _, st = states(world)
    h = make_gpu(world)
    h.set_state(st)
    got = h.get_state()
    assert_state_equal(got, st, world.land)
    assert np.array_equal(got.nplants[world.land], st.nplants[world.land])
    # non-land cells are never written
    assert np.all(got.h2osoi_liq[~world.land] == 0)
    h.close()


# ---- exact mode == host twin of the kernel source: bit-exact ------------------------


def test_exact_single_substep_bitexact_vs_twin(world, forcing30):
    """K1 (one HYDROLOGY call, NISURF=48 => dt=1800 s) from randomised states that
    visit every water-table position; all outputs of the fine-grained entry."""
    _, st = states(world)
    land = world.land
    h = make_gpu(world, mode=MATH_EXACT)
    h.set_state(st)
    out = h.hydrology_step(day_slice(forcing30, 0))
    got = h.get_state()
    tw, ex = oracle_py.twin_run(world, st, {k: v[:1] for k, v in forcing30.items()}, 48,
                                synth.ZI_DRIVER, do_grow=False, math="exact", nsteps=1)
    assert_state_equal(got, tw, land, fields=("h2osoi_liq", "smp", "zwt", "wa"))
    assert np.array_equal(out["jwt"][land], ex["jwt"])
    assert np.array_equal(out["theta"][land], ex["theta"])
    for k in ("qflx_tran_veg_col", "qflx_evap_grnd", "w_imbalance"):
        a, b = out[k][land], ex[k]
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all(), k
    assert out["fault"] == int(np.bitwise_or.reduce(ex["fault"]))
    jwt_in = (st.zwt[land][:, None] > synth.ZI_DRIVER[None, 1:9] / np.float32(1000.0)).sum(axis=1)
    assert (np.bincount(jwt_in, minlength=9) > 0).all()
    h.close()


def test_exact_grow_day_bitexact_vs_twin(world, forcing30):
    _, st = states(world)
    land = world.land
    h = make_gpu(world, mode=MATH_EXACT)
    h.set_state(st)
    out = h.grow_day(np.ascontiguousarray(forcing30["tas"][0]))
    got = h.get_state()
    # GROW alone on the twin: a "day" with zero sub-steps
    tw, ex = oracle_py.twin_run(world, st, {k: v[:1] for k, v in forcing30.items()}, 48,
                                synth.ZI_DRIVER, do_grow=True, math="exact", nsteps=0)
    assert_state_equal(got, tw, land, fields=("lai", "lai_litter", "plant_mass", "plant_foliage_mass",
                                              "plant_length", "rdepth", "rootr_col"))
    assert np.array_equal(out["npp"][land], ex["npp"][0])
    assert np.array_equal(out["w_i"][land], ex["w_i"][0])
    assert np.array_equal(out["fT"][land], ex["fT"][0])
    h.close()


@pytest.mark.parametrize("which", ["init", "random"])
def test_exact_run_days_bitexact_vs_twin(world, forcing30, which):
    """K3 (fused day kernel) through h9_run_days: 10 days x 48 sub-steps + GROW."""
    st_init, st_rand = states(world)
    st = st_init if which == "init" else st_rand
    nd = 10
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing30.items()}
    h = make_gpu(world, mode=MATH_EXACT)
    h.set_state(st)
    rc = h.run_days(np.ones(nd, np.int32), f)
    got = h.get_state()
    tw, ex = oracle_py.twin_run(world, st, f, 48, synth.ZI_DRIVER, math="exact")
    land = world.land
    ok = land.copy()
    ok[land] = ex["fault"] == 0
    assert rc == int(np.bitwise_or.reduce(ex["fault"]))
    assert h.get_fault().n_faulted == int((ex["fault"] != 0).sum())
    assert_state_equal(got, tw, ok, what=f"{which}: ")
    h.close()


# ---- exact mode vs the oracle: stated tolerance ---------------------------------------


@pytest.mark.parametrize("which", ["init", "random"])
def test_exact_vs_oracle_30_days(world, forcing30, which):
    """Tolerance (30 days = 1440 sub-steps): 99.9 % of soil-water values within 1e-4
    relative; every value within 5e-3 relative + 0.02 mm; zwt within 2 mm; annual
    means within 1e-3 relative (+ small absolute floors)."""
    st_init, st_rand = states(world)
    st = st_init if which == "init" else st_rand
    nd = 30
    yi = np.ones(nd, np.int32)
    o = make_oracle(world)
    o.set_state(st)
    orc = o.run_days(yi, forcing30)
    ref = o.get_state()
    h = make_gpu(world, mode=MATH_EXACT)
    h.set_state(st)
    rc = h.run_days(yi, forcing30)
    got = h.get_state()
    assert rc == orc == 0
    land = world.land
    rel = relerr(got.h2osoi_liq[land], ref.h2osoi_liq[land])
    record(f"exact_vs_oracle_{which}_30d", h2o_rel_p999=np.quantile(rel, 0.999), h2o_rel_max=rel.max(),
           zwt_abs_max=np.abs(got.zwt[land] - ref.zwt[land]).max())
    assert np.quantile(rel, 0.999) < 1e-4
    assert_state_close(got, ref, land, rtol=5e-3, atol=0.02, fields=("h2osoi_liq", "wa"))
    assert_state_close(got, ref, land, rtol=0, atol=2e-3, fields=("zwt",))
    assert_state_close(got, ref, land, rtol=2e-3, atol=1e-5,
                       fields=("lai", "lai_litter", "plant_mass", "plant_foliage_mass", "rootr_col"))
    ga, oa = h.get_annual(1), o.get_annual(1)
    for k, at in (("npp", 1e-3), ("plant_mass", 1e-4), ("rnf", 1e-7), ("theta_total", 0.05), ("theta", 1e-5)):
        err = np.abs(ga[k][land].astype(np.float64) - oa[k][land]) - (at + 2e-3 * np.abs(oa[k][land]))
        assert (err <= 0).all(), (k, err.max())
    assert np.all(ga["evap"][land] == 0)
    assert np.isnan(ga["npp"][~land]).all() and np.all(ga["theta_total"][~land] == 0)
    h.close()


# ---- fast mode vs the oracle, against the rounding-noise floor -----------------------------


@pytest.mark.parametrize("block", FAST_KERNELS, ids=FAST_KERNEL_IDS)
@pytest.mark.parametrize("which", ["init", "random"])
def test_fast_vs_oracle_30_days(world, forcing30, which, block):
    """H9_MATH_FAST (MUFU pow/exp/rcp, FMA contraction), both stepping kernels (thread per
    cell; two lanes per cell with the two-sided tridiagonal solve).  Stated tolerance for 30 days:
    median relative error of soil water < 2e-5, 99.9 % within 5e-3, every value within
    5e-2 relative + 0.5 mm, and the worst error no more than 10x the FP32 rounding-noise
    floor measured by the float-vs-double oracle on the same case."""
    st_init, st_rand = states(world)
    st = st_init if which == "init" else st_rand
    nd = 30
    yi = np.ones(nd, np.int32)
    land = world.land
    res = {}
    for kind in ("f32", "f64"):
        o = make_oracle(world, kind=kind)
        o.set_state(st)
        assert o.run_days(yi, forcing30) == 0
        res[kind] = o.get_state()
    h = make_gpu(world, mode=MATH_FAST, block=block)
    assert ("pair" in h.kernel_variant()) == (block == TWO_LANES)
    h.set_state(st)
    assert h.run_days(yi, forcing30) == 0
    got = h.get_state()
    rel = relerr(got.h2osoi_liq[land], res["f32"].h2osoi_liq[land])
    noise = relerr(res["f32"].h2osoi_liq[land], res["f64"].h2osoi_liq[land])
    record(f"fast_vs_oracle_{which}_30d_{FAST_KERNEL_IDS[FAST_KERNELS.index(block)]}", h2o_rel_p50=np.median(rel), h2o_rel_p999=np.quantile(rel, 0.999),
           h2o_rel_max=rel.max(), noise_p50=np.median(noise), noise_p999=np.quantile(noise, 0.999),
           noise_max=noise.max(), zwt_abs_max=np.abs(got.zwt[land] - res["f32"].zwt[land]).max(),
           lai_rel_max=relerr(got.lai[land], res["f32"].lai[land]).max())
    assert np.median(rel) < 2e-5
    assert np.quantile(rel, 0.999) < 5e-3
    assert_state_close(got, res["f32"], land, rtol=5e-2, atol=0.5, fields=("h2osoi_liq", "wa"))
    assert rel.max() < 10 * max(noise.max(), 1e-3)
    assert np.abs(got.zwt[land] - res["f32"].zwt[land]).max() < 2e-2
    assert_state_close(got, res["f32"], land, rtol=2e-2, atol=1e-4, fields=("lai", "plant_mass"))
    h.close()


@pytest.mark.parametrize("block", FAST_KERNELS, ids=FAST_KERNEL_IDS)
def test_fast_single_substep_vs_oracle(world, forcing30, block):
    """One sub-step from randomised (deliberately extreme) states, fast mode: 99 % of soil
    water within 1e-3 relative, every value within 2e-2 relative + 0.05 mm, and the worst
    error within 10x the float-vs-double rounding noise of the same step."""
    _, st = states(world)
    land = world.land
    o = make_oracle(world)
    o.set_state(st)
    oo = o.hydrology_step(day_slice(forcing30, 0))
    ref = o.get_state()
    h = make_gpu(world, mode=MATH_FAST, block=block)
    h.set_state(st)
    go = h.hydrology_step(day_slice(forcing30, 0))
    got = h.get_state()
    o64 = make_oracle(world, kind="f64")
    o64.set_state(st)
    o64.hydrology_step(day_slice(forcing30, 0))
    ok = land & (np.abs(oo["w_imbalance"]) <= 0.05) & (np.abs(go["w_imbalance"]) <= 0.05)
    rel = relerr(got.h2osoi_liq[ok], ref.h2osoi_liq[ok])
    noise = relerr(ref.h2osoi_liq[ok], o64.get_state().h2osoi_liq[ok])
    record(f"fast_single_step_{FAST_KERNEL_IDS[FAST_KERNELS.index(block)]}", h2o_rel_p50=np.median(rel), h2o_rel_p99=np.quantile(rel, 0.99),
           h2o_rel_max=rel.max(), noise_p99=np.quantile(noise, 0.99), noise_max=noise.max(),
           n_cells=int(ok.sum()), jwt_agree=(go["jwt"][ok] == oo["jwt"][ok]).mean())
    assert np.median(rel) < 1e-5
    assert np.quantile(rel, 0.99) < 1e-3
    assert_state_close(got, ref, ok, rtol=2e-2, atol=0.05, fields=("h2osoi_liq", "wa"))
    assert rel.max() < 10 * max(noise.max(), 1e-3)
    assert (go["jwt"][ok] == oo["jwt"][ok]).mean() > 0.995
    h.close()


# ---- pipeline equivalences: bit-exact --------------------------------------------------------


@pytest.mark.parametrize("mode", [MATH_EXACT, MATH_FAST])
def test_run_days_equals_device_resident_path(world, forcing30, mode):
    """h9_run_days (tiled H2D + pack + step) == h9_pack_forcing + h9_run_days_device,
    for any tile size and block size, including a year change inside the batch."""
    _, st = states(world)
    nd = 13
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing30.items()}
    yi = np.concatenate([np.full(6, 1, np.int32), np.full(nd - 6, 2, np.int32)])
    results = []
    # every launch shape and both builds of the thread-per-cell kernel (all registers: the
    # small-shard step with its straight-line tails; 128 registers: the throughput step)
    for tile, block, device_path in ((8, 64, False), (3, 32, False), (5, 128, False), (8, 64, True),
                                     (8, THREAD_128REG, False), (4, 1032, True)):
        h = make_gpu(world, mode=mode, nyr=2)
        h.set_tuning(tile, block)
        h.set_state(st)
        if device_path:
            p, ds, ps = h.pack_forcing(f, nd)
            h.run_days_device(yi, p, ds, ps)
        else:
            h.run_days(yi, f)
        results.append((h.get_state(), h.get_annual(1), h.get_annual(2)))
        h.close()
    s0, a1, a2 = results[0]
    for s, b1, b2 in results[1:]:
        assert_state_equal(s, s0, world.land)
        for k in a1:
            assert np.array_equal(a1[k], b1[k], equal_nan=True), k
            assert np.array_equal(a2[k], b2[k], equal_nan=True), k


def test_split_calls_equal_one_call(world, forcing30):
    st_init, _ = states(world)
    nd = 12
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing30.items()}
    a = make_gpu(world, mode=MATH_FAST)
    a.set_state(st_init)
    a.run_days(np.ones(nd, np.int32), f)
    b = make_gpu(world, mode=MATH_FAST)
    b.set_state(st_init)
    for d0 in range(0, nd, 5):
        d1 = min(nd, d0 + 5)
        b.run_days(np.ones(d1 - d0, np.int32), {k: np.ascontiguousarray(v[d0:d1]) for k, v in f.items()})
    assert_state_equal(a.get_state(), b.get_state(), world.land)
    for k, v in a.get_annual(1).items():
        assert np.array_equal(v, b.get_annual(1)[k], equal_nan=True), k
    a.close()
    b.close()


def test_pack_forcing_layout(world, forcing30):
    """K4: (lon_c,lat_c,ndays) host arrays -> [day][7][ncs] compact device layout."""
    import torch
    from hybrid9_b200.distributed import device_tensor
    nd = 9
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing30.items()}
    h = make_gpu(world)
    h.set_tuning(4, 0)
    p, ds, ps = h.pack_forcing(f, nd)
    n = h.num_land
    assert ds == 7 * ps and ps % 128 == 0 and ps >= n
    t = device_tensor(p, (nd, 7, ps), torch.float32, torch.device("cuda", 0)).cpu().numpy()
    yy, xx = np.nonzero(world.land)
    for j, k in enumerate(oracle_py.FORCING):
        assert np.array_equal(t[:, j, :n], f[k][:, yy, xx]), k
    h.close()


def test_pageable_and_pinned_host_forcing_agree(world, forcing30):
    from hybrid9_b200.host import pinned_empty
    st_init, _ = states(world)
    nd = 9
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing30.items()}
    fp = {k: pinned_empty(v.shape) for k, v in f.items()}
    for k in f:
        fp[k][...] = f[k]
    out = []
    for ff in (f, fp):
        h = make_gpu(world, mode=MATH_FAST)
        h.set_tuning(4, 0)
        h.set_state(st_init)
        h.run_days(np.ones(nd, np.int32), ff)
        out.append(h.get_state())
        h.close()
    assert_state_equal(out[0], out[1], world.land)


# ---- faults: the reference's STOP conditions ----------------------------------------------------


def test_water_imbalance_fault_is_reported_like_the_reference(world, forcing30):
    """NISURF=1 (dt = 86400 s, a configuration the reference's notes mention) from
    randomised states makes some cells trip |w1-w0| > 0.1 mm (HYDROLOGY.f90:1244).  The
    GPU returns the fault bits and the first fault's cell / day / sub-step / code /
    imbalance: bit-exact against the host twin, and the same record as the oracle."""
    _, st = states(world)
    land = world.land
    f = {k: np.ascontiguousarray(v[:1]) for k, v in forcing30.items()}
    o = make_oracle(world, nisurf=1)
    o.set_state(st)
    orc = o.run_days(np.ones(1, np.int32), f)
    of = o.get_fault()
    assert orc & 8 and of["n_faulted"] > 0, "test state no longer trips the reference's STOP condition"
    tw, ex = oracle_py.twin_run(world, st, f, 1, synth.ZI_DRIVER, math="exact")
    h = make_gpu(world, nisurf=1, mode=MATH_EXACT)
    h.set_state(st)
    rc = h.run_days(np.ones(1, np.int32), f)
    gf = h.get_fault()
    first = int(np.flatnonzero(ex["fault"])[0])
    yy, xx = np.nonzero(land)
    assert rc == int(np.bitwise_or.reduce(ex["fault"])) and gf.any == rc
    assert gf.n_faulted == int((ex["fault"] != 0).sum())
    assert (gf.x, gf.y, gf.day, gf.substep) == (int(xx[first]) + 1, int(yy[first]) + 1, 1, 1)
    assert gf.code == int(ex["fault"][first])
    assert np.float32(gf.imbalance) == ex["w_imbalance"][first]
    # the oracle (libm pow) sees the same first fault; counts may differ by cells at the threshold
    assert (gf.x, gf.y, gf.day, gf.substep, gf.code) == (of["x"], of["y"], of["day"], of["substep"], of["code"])
    assert abs(gf.n_faulted - of["n_faulted"]) <= max(2, of["n_faulted"] // 10)
    assert np.isclose(gf.imbalance, of["imbalance"], rtol=1e-2, atol=1e-3)
    h.clear_fault()
    assert h.get_fault().any == 0
    h.close()


# ---- annual diagnostics and the device-side view ---------------------------------------------------


def test_annual_device_view_and_budget(world, forcing30):
    import torch
    from hybrid9_b200.distributed import device_tensor
    st_init, _ = states(world)
    nd = 8
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing30.items()}
    h = make_gpu(world, mode=MATH_FAST)
    h.set_state(st_init)
    h.run_days(np.ones(nd, np.int32), f)
    ann = h.get_annual(1)
    p, stride, bp = h.annual_device(1)
    h.synchronize()
    dev = torch.device("cuda", 0)
    m = device_tensor(p, (13, stride), torch.float32, dev).cpu().numpy()
    b = device_tensor(bp, (8,), torch.float64, dev).cpu().numpy()
    land = world.land
    n = h.num_land
    assert np.array_equal(m[0, :n], ann["npp"][land]) and np.array_equal(m[2, :n], ann["rnf"][land])
    assert np.array_equal(m[5:13, :n].T, ann["theta"][land])
    st = h.get_state()
    assert np.isclose(b[0], st.h2osoi_liq[land].astype(np.float64).sum(), rtol=1e-12)
    assert np.isclose(b[1], st.wa[land].astype(np.float64).sum(), rtol=1e-12)
    assert np.isclose(b[2], ann["rnf"][land].astype(np.float64).sum(), rtol=1e-12)
    assert np.isclose(b[3], ann["npp"][land].astype(np.float64).sum(), rtol=1e-12)
    assert b[5] == n and b[7] == 0
    h.close()


# ---- degenerate blocks --------------------------------------------------------------------------------


def test_empty_block_and_single_cell(world):
    yy, xx = np.nonzero(~world.land)
    e = world.window(int(xx[0]) + 1, int(yy[0]) + 1, 3, 2)
    e.soil_tex[...] = 0
    h = make_gpu(e)
    assert h.num_land == 0
    f = synth.make_forcing(e, 2, seed=1, land_only=False)
    h.set_state(init_state(e.soil_tex, e.theta_s, synth.ZI_DRIVER))
    assert h.run_days(np.ones(2, np.int32), f) == 0
    assert np.isnan(h.get_annual(1)["npp"]).all()
    h.close()
    yy, xx = np.nonzero(world.land)
    s = world.window(int(xx[5]) + 1, int(yy[5]) + 1, 1, 1)
    f = synth.make_forcing(s, 20, seed=1)
    st = init_state(s.soil_tex, s.theta_s, synth.ZI_DRIVER)
    h = make_gpu(s, mode=MATH_EXACT)
    h.set_state(st)
    assert h.num_land == 1 and h.run_days(np.ones(20, np.int32), f) == 0
    tw, _ = oracle_py.twin_run(s, st, f, 48, synth.ZI_DRIVER, math="exact")
    assert_state_equal(h.get_state(), tw, s.land)
    h.close()


def test_missing_lambda_cells_do_not_poison_neighbours():
    """bsw = 1e8 (INIT.f90:624-628, G28): such cells may overflow and fault, all other
    cells are unaffected (cells are independent)."""
    w = synth.make_world(nx=72, ny=36, seed=5, n_missing_lambda=4)
    w0 = synth.make_world(nx=72, ny=36, seed=5)
    bad = (w.bsw[..., 0] > 1e7) & w.land
    assert bad.sum() == 4
    f = synth.make_forcing(w, 3, seed=2)
    st = init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER)
    out = []
    for ww in (w, w0):
        h = make_gpu(ww, mode=MATH_FAST)
        h.set_state(st)
        h.run_days(np.ones(3, np.int32), f)
        out.append(h.get_state())
        h.close()
    good = w.land & ~bad
    assert_state_equal(out[0], out[1], good)


def test_real_evaporation_option(world, forcing30):
    """H9_OPT_REAL_EVAP: by default axy_evap == 0 like the reference (evap_sum is never
    accumulated, HYBRID9.f90:137,276); with the option it is the annual mean ET flux."""
    from hybrid9_b200.host import OPT_REAL_EVAP
    st_init, _ = states(world)
    nd = 10
    f = {k: np.ascontiguousarray(v[:nd]) for k, v in forcing30.items()}
    yi = np.ones(nd, np.int32)
    o = make_oracle(world)
    o.set_real_evap(True)
    o.set_state(st_init)
    o.run_days(yi, f)
    ref = o.get_annual(1)["evap"]
    land = world.land
    assert (ref[land] > 0).mean() > 0.5
    for mode, rtol in ((MATH_EXACT, 1e-4), (MATH_FAST, 5e-3)):
        h = make_gpu(world, mode=mode)
        h.set_option(OPT_REAL_EVAP, 1)
        h.set_state(st_init)
        assert h.run_days(yi, f) == 0
        got = h.get_annual(1)["evap"]
        err = np.abs(got[land].astype(np.float64) - ref[land]) - (1e-9 + rtol * np.abs(ref[land]))
        assert (err <= 0).all(), (mode, err.max())
        h.set_option(OPT_REAL_EVAP, 0)
        h.close()
