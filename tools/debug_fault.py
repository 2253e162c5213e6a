"""Reproduce a fast-mode fault on one cell and dump the state just before it."""
import sys, os, pickle
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hybrid9_b200 import H9, MATH_FAST, MATH_EXACT, synth
from hybrid9_b200.state import init_state
x, y = int(sys.argv[1]), int(sys.argv[2])
mode = MATH_FAST if (len(sys.argv) < 4 or sys.argv[3] == "fast") else MATH_EXACT
w = synth.make_world()
forcing = synth.make_forcing(w, 365, seed=9)
s = w.window(x, y, 1, 1)
f = {k: np.ascontiguousarray(v[:, y-1:y, x-1:x]) for k, v in forcing.items()}
h = H9(0); h.configure(1, 1, 48, synth.ZI_DRIVER, nyr=2); h.set_math(mode)
h.set_soil(s.soil_tex, s.theta_s, s.hksat, s.bsw, s.psi_s, s.fmax)
h.set_state(init_state(s.soil_tex, s.theta_s, synth.ZI_DRIVER), with_smp=False)
out = {"world": s, "forcing": f}
day_abs = 0
for yr in range(30):
    for d in range(365):
        pre = h.get_state()
        fd = {k: np.ascontiguousarray(v[d:d+1]) for k, v in f.items()}
        rc = h.run_days(np.full(1, yr % 2 + 1, np.int32), fd)
        day_abs += 1
        if rc:
            print("fault at year", yr + 1, "day", d + 1, "abs", day_abs, h.get_fault())
            # replay the day sub-step by sub-step from `pre`
            h2 = H9(0); h2.configure(1, 1, 48, synth.ZI_DRIVER, nyr=2); h2.set_math(mode)
            h2.set_soil(s.soil_tex, s.theta_s, s.hksat, s.bsw, s.psi_s, s.fmax)
            h2.set_state(pre)
            f1 = {k: np.ascontiguousarray(v[d]) for k, v in f.items()}
            for ns in range(48):
                p2 = h2.get_state()
                o = h2.hydrology_step(f1)
                print(ns + 1, "imb", o["w_imbalance"].ravel(), "jwt", o["jwt"].ravel(), "zwt", h2.get_state().zwt.ravel())
                if o["fault"]:
                    out.update(pre_state=p2, day=d, substep=ns + 1, post_state=h2.get_state(), step_out=o)
                    break
            pickle.dump(out, open(os.path.join(ROOT, "gpurun_out", "fault_dump.pkl"), "wb"))
            sys.exit(0)
print("no fault in 30 years")
