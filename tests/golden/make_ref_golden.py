"""Generate tests/golden/h9_ref_golden_v1.npz: outputs of the REFERENCE'S OWN code.

Unlike h9_golden_v1.npz (a frozen oracle output), these vectors come from
oracle/_ref/libh9ref.so: HYDROLOGY.f90, GROW.f90 and the loop nest of HYBRID9.f90 translated
statement for statement by oracle/f2cpp.py from /root/reference/SOURCE and compiled with g++
(strict IEEE, glibc powf/expf/logf).  /root/reference does not exist on the GPU box, so the
vectors are committed; tests/test_ref_golden.py holds the oracle to them bit for bit (CPU) and
the GPU within stated tolerances.

Cases
  init / random   24 x 12 block, 12 days over 2 year slots, NISURF 48, from the INIT state and
                  from a branch-coverage state (water table in every layer); smp per cell
  leak            the same block from INIT with the reference's ONE shared smp scratch vector
                  (SHARED.f90:198), verbatim loop nest HYBRID9.f90:120-295
  year            BASELINE.json configs[0], the correctness anchor: a 4 x 2 block, one whole
                  calendar year (1901, 365 days x 48 sub-steps) through HYBRID9.f90:103-295

Regenerate (needs /root/reference): python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_py  # noqa: E402
from helpers import STATE_FIELDS  # noqa: E402
from hybrid9_b200 import calendar, synth  # noqa: E402

NX, NY, SEED, NDAYS, NISURF = 24, 12, 17, 12, 48
YNX, YNY, YSEED = 4, 2, 3
FIELDS = STATE_FIELDS + ("nplants",)
PATH = os.path.join(HERE, "h9_ref_golden_v1.npz")


def small_world():
    return synth.make_world(nx=NX, ny=NY, seed=SEED, n_class13=2, n_zero_theta=2)


def year_world():
    return synth.make_world(nx=YNX, ny=YNY, seed=YSEED)


def year_index():
    return np.concatenate([np.full(7, 1, np.int32), np.full(NDAYS - 7, 2, np.int32)])


def year_days():
    return calendar.time_boy(1902) - calendar.time_boy(1901)


def build():
    out = {}
    w = small_world()
    f = synth.make_forcing(w, NDAYS, seed=SEED)
    yi = year_index()
    for tag in ("init", "random", "leak"):
        r = ref_py.make_ref(w, nisurf=NISURF, nyr=2, per_cell_smp=(tag != "leak"))
        r.init_state()
        st0 = r.get_state()
        if tag == "random":
            st0 = synth.randomize_state(w, st0, seed=SEED)
        r.set_state(st0)
        out[f"{tag}_rc"] = np.int32(r.run_days(yi, f))
        st1 = r.get_state()
        for n in FIELDS:
            out[f"{tag}_in_{n}"] = getattr(st0, n)
            out[f"{tag}_out_{n}"] = getattr(st1, n)
        for iy in (1, 2):
            for k, v in r.get_annual(iy).items():
                out[f"{tag}_axy{iy}_{k}"] = v
    # one HYDROLOGY call and one GROW call from the branch-coverage state, all outputs
    r = ref_py.make_ref(w, nisurf=NISURF)
    r.init_state()
    st0 = synth.randomize_state(w, r.get_state(), seed=SEED + 1)
    r.set_state(st0)
    f0 = {k: np.ascontiguousarray(v[0]) for k, v in f.items()}
    step = r.hydrology_step(f0)
    for n in FIELDS:
        out[f"step_in_{n}"] = getattr(st0, n)
        out[f"step_out_{n}"] = getattr(r.get_state(), n)
    for k in ("theta", "qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc", "w_imbalance", "jwt"):
        out[f"step_{k}"] = step[k]
    out["step_fault"] = np.int32(step["fault"])
    g = r.grow_day(f0["tas"])
    for k, v in g.items():
        out[f"grow_{k}"] = v
    for n in FIELDS:
        out[f"grow_out_{n}"] = getattr(r.get_state(), n)
    # configs[0]: one calendar year of a handful of cells on the reference calendar
    wy = year_world()
    nd = year_days()
    fy = synth.make_forcing(wy, calendar.decade_days(1), seed=YSEED)
    fy1 = {k: np.ascontiguousarray(v[:nd]) for k, v in fy.items()}
    r = ref_py.make_ref(wy, nisurf=NISURF, nyr=1, per_cell_smp=True)
    r.init_state()
    out["year_rc"] = np.int32(r.run_days(np.ones(nd, np.int32), fy1))
    st1 = r.get_state()
    for n in FIELDS:
        out[f"year_out_{n}"] = getattr(st1, n)
    for k, v in r.get_annual(1).items():
        out[f"year_axy1_{k}"] = v
    return out


if __name__ == "__main__":
    np.savez_compressed(PATH, **build())
    print("wrote", PATH, os.path.getsize(PATH), "bytes")
