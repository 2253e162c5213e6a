"""Diagnostic: share of land cells whose water table is inside the soil column (jwt < 8) and the
share of 32-cell warps that hold at least one such cell, year by year, on the bench workload."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid9_b200 import H9, MATH_FAST, synth
from hybrid9_b200.state import init_state

w = synth.make_world()
f = synth.make_forcing(w, 365, seed=9)
h = H9(0); h.configure(w.nx, w.ny, 48, synth.ZI_DRIVER, nyr=1); h.set_math(MATH_FAST)
h.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax)
h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), with_smp=False)
land = w.land
out = []
for yr in range(1, int(sys.argv[1]) + 1 if len(sys.argv) > 1 else 9):
    t = time.time(); h.run_days(np.ones(365, np.int32), f); h.synchronize(); dt = time.time() - t
    z = h.get_state().zwt[land]
    sh = z <= np.float32(2.296)
    n = sh.size // 32 * 32
    anyw = sh[:n].reshape(-1, 32).any(axis=1).mean()
    srt = np.sort(sh[:n])[::-1].reshape(-1, 32).any(axis=1).mean()
    out.append(dict(year=yr, shallow=float(sh.mean()), warps_with_shallow=float(anyw), after_grouping=float(srt), wall_s=dt))
    print(out[-1], flush=True)
json.dump(out, open("gpurun_out/regime_stats.json", "w"), indent=1)
