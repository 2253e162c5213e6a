"""The two-lanes-per-cell stepping kernel (h9_kernels_pair.cu, small shards) against the
thread-per-cell kernel and the oracle, through the C ABI (`-m gpu`).

Both kernels are H9_MATH_FAST: same MUFU pow/exp/rcp, same per-layer arithmetic; the pair
kernel solves the nine tridiagonal equations of HYDROLOGY.f90:661-831 from both ends and sums
the column in a different order.  Indexing (land index, jwt, fault cell/day/sub-step) must
agree exactly; values at rounding level, stated per test.  The oracle gates of
tests/test_gpu_parity.py run on both kernels (parametrised there)."""
import numpy as np
import pytest

from helpers import (THREAD_PER_CELL, TWO_LANES, assert_state_close, assert_state_equal, day_slice,
                     make_gpu, make_oracle)
from hybrid9_b200 import MATH_FAST, synth
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    return synth.make_world(nx=144, ny=72, seed=5)


@pytest.fixture(scope="module")
def forcing(world):
    return synth.make_forcing(world, 12, seed=3)


def rel(a, b, floor=1e-3):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def test_automatic_choice_by_shard_size(world):
    h = make_gpu(world, mode=MATH_FAST)
    assert h.num_land < 9472 and h.kernel_variant() == "h9::days_kernel_pair<128>"
    h.set_tuning(0, THREAD_PER_CELL)
    assert h.kernel_variant() == "h9::days_kernel_fast<64,1>"
    h.set_tuning(0, 1064)
    assert "pair" in h.kernel_variant()
    h.close()


def test_single_substep_pair_vs_thread_per_cell(world, forcing):
    """K1 from randomised states with the water table in every layer (both Drainage regimes,
    cascade and dryness repair included): jwt identical, soil water within 2e-6 relative in the
    median and 1e-4 for 99.9 %, fluxes within 1e-4 relative, same fault bits."""
    st = synth.randomize_state(world, init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), seed=11)
    land = world.land
    out, state = {}, {}
    for name, block in (("thread", THREAD_PER_CELL), ("pair", TWO_LANES)):
        h = make_gpu(world, mode=MATH_FAST, block=block)
        h.set_state(st)
        out[name] = h.hydrology_step(day_slice(forcing, 0))
        state[name] = h.get_state()
        h.close()
    a, b = out["pair"], out["thread"]
    assert a["fault"] == b["fault"]
    ok = land & (np.abs(b["w_imbalance"]) <= 0.05)
    jwt_in = (st.zwt[land][:, None] > synth.ZI_DRIVER[None, 1:9] / np.float32(1000.0)).sum(axis=1)
    assert (np.bincount(jwt_in, minlength=9) > 0).all()
    assert (a["jwt"][ok] == b["jwt"][ok]).mean() > 0.999
    r = rel(state["pair"].h2osoi_liq[ok], state["thread"].h2osoi_liq[ok])
    assert np.median(r) < 2e-6 and np.quantile(r, 0.999) < 1e-4, (np.median(r), np.quantile(r, 0.999), r.max())
    assert_state_close(state["pair"], state["thread"], ok, rtol=5e-3, atol=0.02, fields=("h2osoi_liq", "wa"))
    assert_state_close(state["pair"], state["thread"], ok, rtol=0, atol=1e-3, fields=("zwt",))
    for k in ("qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc"):
        e = np.abs(a[k][ok].astype(np.float64) - b[k][ok]) - (1e-12 + 1e-4 * np.abs(b[k][ok]))
        assert (e <= 0).all(), (k, e.max())
    r = rel(a["theta"][ok], b["theta"][ok])
    assert np.quantile(r, 0.999) < 1e-4
    # smp is per-layer arithmetic only: the two kernels give the same bits
    assert np.array_equal(state["pair"].smp[ok], state["thread"].smp[ok])


@pytest.mark.parametrize("which", ["init", "random"])
def test_fused_days_pair_vs_thread_per_cell(world, forcing, which):
    """K3: 12 days x 48 sub-steps + GROW + annual means, two year slots."""
    st0 = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    st = st0 if which == "init" else synth.randomize_state(world, st0, seed=11)
    nd = 12
    yi = np.concatenate([np.full(5, 1, np.int32), np.full(nd - 5, 2, np.int32)])
    land = world.land
    res = {}
    for name, block in (("thread", THREAD_PER_CELL), ("pair", TWO_LANES)):
        h = make_gpu(world, mode=MATH_FAST, nyr=2, block=block)
        h.set_state(st)
        rc = h.run_days(yi, forcing)
        res[name] = (rc, h.get_state(), h.get_annual(1), h.get_annual(2), h.get_fault())
        h.close()
    (rca, sa, a1, a2, fa), (rcb, sb, b1, b2, fb) = res["pair"], res["thread"]
    assert rca == rcb and fa.n_faulted == fb.n_faulted
    r = rel(sa.h2osoi_liq[land], sb.h2osoi_liq[land])
    assert np.median(r) < 5e-6 and np.quantile(r, 0.999) < 5e-3, (np.median(r), np.quantile(r, 0.999))
    assert np.quantile(np.abs(sa.zwt[land] - sb.zwt[land]), 0.999) < 2e-3
    assert_state_close(sa, sb, land, rtol=2e-3, atol=1e-5,
                       fields=("lai", "lai_litter", "plant_mass", "plant_foliage_mass", "rootr_col"))
    for x, y in ((a1, b1), (a2, b2)):
        for k in x:
            assert np.array_equal(np.isnan(x[k]), np.isnan(y[k])), k       # fills of INIT.f90:402-414
            assert np.array_equal(x[k][~land], y[k][~land], equal_nan=True), k
        for k, at in (("npp", 1e-3), ("plant_mass", 1e-4), ("rnf", 1e-7), ("theta_total", 0.05), ("theta", 1e-5)):
            e = np.abs(x[k][land].astype(np.float64) - y[k][land]) - (at + 2e-3 * np.abs(y[k][land]))
            assert (e <= 0).all(), (k, e.max())


def test_pair_is_deterministic_and_split_invariant(world, forcing):
    """Same bits run to run, for split calls, and for the device-resident entry."""
    st = init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER)
    nd = 12
    yi = np.ones(nd, np.int32)
    runs = []
    for split, device_path in ((False, False), (True, False), (False, True)):
        h = make_gpu(world, mode=MATH_FAST, block=TWO_LANES)
        h.set_state(st)
        if device_path:
            p, ds, ps = h.pack_forcing(forcing, nd)
            h.run_days_device(yi, p, ds, ps)
        elif split:
            for d0 in range(0, nd, 5):
                d1 = min(nd, d0 + 5)
                h.run_days(yi[d0:d1], {k: np.ascontiguousarray(v[d0:d1]) for k, v in forcing.items()})
        else:
            h.run_days(yi, forcing)
        runs.append((h.get_state(), h.get_annual(1)))
        h.close()
    for s, a in runs[1:]:
        assert_state_equal(s, runs[0][0], world.land)
        for k in a:
            assert np.array_equal(a[k], runs[0][1][k], equal_nan=True), k


def test_pair_odd_cell_counts_and_tiny_blocks():
    """Shard sizes that leave half-filled warps and blocks: 1 cell, 3 cells, 65 cells."""
    for nx, ny, n_land in ((4, 3, 1), (5, 4, 3), (16, 12, 65)):
        w = synth.make_world(nx=nx, ny=ny, n_land=n_land, seed=2)
        f = synth.make_forcing(w, 3, seed=4)
        st = init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER)
        res = []
        for block in (THREAD_PER_CELL, TWO_LANES):
            h = make_gpu(w, mode=MATH_FAST, block=block)
            h.set_state(st)
            assert h.run_days(np.ones(3, np.int32), f) == 0
            res.append(h.get_state())
            h.close()
        land = w.land
        assert int(land.sum()) == n_land
        r = rel(res[1].h2osoi_liq[land], res[0].h2osoi_liq[land])
        assert r.max() < 1e-3, (n_land, r.max())
        assert np.array_equal(res[1].h2osoi_liq[~land], res[0].h2osoi_liq[~land])


def test_pair_stop_record_matches_the_oracle(world, forcing):
    """NISURF = 1 from randomised states trips the reference's |w1-w0| > 0.1 STOP
    (HYDROLOGY.f90:1244-1274): same first cell, day, sub-step and fault bits as the oracle and
    as the thread-per-cell kernel."""
    st = synth.randomize_state(world, init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), seed=11)
    f = {k: np.ascontiguousarray(v[:1]) for k, v in forcing.items()}
    yi = np.ones(1, np.int32)
    o = make_oracle(world, nisurf=1)
    o.set_state(st)
    orc = o.run_days(yi, f)
    of = o.get_fault()
    assert orc & 8 and of["n_faulted"] > 0
    rec = {}
    for name, block in (("thread", THREAD_PER_CELL), ("pair", TWO_LANES)):
        h = make_gpu(world, nisurf=1, mode=MATH_FAST, block=block)
        h.set_state(st)
        rc = h.run_days(yi, f)
        rec[name] = (rc, h.get_fault())
        h.close()
    rc, gf = rec["pair"]
    assert rc == rec["thread"][0] and rc & 8
    assert (gf.x, gf.y, gf.day, gf.substep, gf.code) == (of["x"], of["y"], of["day"], of["substep"], of["code"])
    tf = rec["thread"][1]
    assert (gf.x, gf.y, gf.day, gf.substep, gf.code) == (tf.x, tf.y, tf.day, tf.substep, tf.code)
    assert abs(gf.n_faulted - of["n_faulted"]) <= max(2, of["n_faulted"] // 10)
    # a day-long step is far outside the solver's design range: the two-sided elimination and the
    # forward sweep round differently there, so only the sign and the STOP itself are compared
    assert abs(gf.imbalance) > 0.1 and np.sign(gf.imbalance) == np.sign(of["imbalance"])
