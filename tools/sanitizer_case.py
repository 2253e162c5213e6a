"""One pass over every kernel of libh9gpu on a tiny grid, for compute-sanitizer
(tools/sanitize.sh): the fused day kernels (exact; fast thread-per-cell in both builds; fast
two-lanes-per-cell), the one-routine entries K1/K2 in every variant, the forcing pack K4 (both
ingest paths), the budget kernel K5, the soil regrid K6, state round trips and the fault
read-back.  Randomised states so that the data-dependent Drainage code, the cascade and the
dryness repair all execute.  Prints the launches it made; the sanitizer prints the verdict."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid9_b200 import H9, MATH_EXACT, MATH_FAST, synth  # noqa: E402
from hybrid9_b200.state import init_state  # noqa: E402


def main():
    w = synth.make_world(nx=36, ny=18, seed=9)
    nd = 2
    f = synth.make_forcing(w, nd, seed=9)
    st0 = init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER)
    st1 = synth.randomize_state(w, st0, seed=11)
    launches = 0
    for mode, block in ((MATH_EXACT, 0), (MATH_FAST, 64), (MATH_FAST, 1128), (MATH_FAST, 4000)):
        for st in (st0, st1):
            h = H9(0)
            h.configure(w.nx, w.ny, 48, synth.ZI_DRIVER, nyr=2)
            h.set_math(mode)
            if block:
                h.set_tuning(0, block)
            h.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax)
            h.set_state(st)
            h.hydrology_step({k: np.ascontiguousarray(v[0]) for k, v in f.items()})   # K1
            h.grow_day(np.ascontiguousarray(f["tas"][0]))                               # K2
            h.run_days(np.array([1, 2], np.int32), f)                                   # K3, gather ingest
            p, ds, ps = h.pack_forcing(f, nd)                                           # K4
            h.run_days_device(np.array([2, 2], np.int32), p, ds, ps)                    # K3, compact forcing
            h.annual_device(2, budget=True)                                             # K5
            h.get_annual(1)
            h.get_state()
            h.get_fault()
            h.clear_fault()
            launches += h.counters()["launches"]
            h.close()
    # K6: soil regrid of a 3 x 2 coarse block
    h = H9(0)
    lon_c, lat_c = 3, 2
    rng = np.random.default_rng(1)
    fine = [np.ascontiguousarray(rng.uniform(0.1, 500.0, (lat_c * 60, lon_c * 60)).astype(np.float32)) for _ in range(4)]
    fine[0][::7, ::5] = -1.0
    outs = [np.zeros((lat_c, lon_c, 8), np.float32) for _ in range(4)]
    h.regrid_soil_layer(lon_c, lat_c, 3, *fine, *outs)
    launches += h.counters()["launches"]
    h.close()
    print(f"sanitizer case done: {launches} kernel launches")


if __name__ == "__main__":
    main()
