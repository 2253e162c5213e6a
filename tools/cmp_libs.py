import os, sys, subprocess, numpy as np, pickle
# run in subprocess per library (library is loaded once per process)
code = r'''
import sys, numpy as np, pickle
sys.path.insert(0, ".")
from hybrid9_b200 import H9, MATH_FAST, synth
from hybrid9_b200.state import init_state
w = synth.make_world(nx=144, ny=72, seed=5)
f = synth.make_forcing(w, 20, seed=3)
st = synth.randomize_state(w, init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), seed=11)
h = H9(0); h.configure(w.nx, w.ny, 48, synth.ZI_DRIVER, nyr=1); h.set_math(MATH_FAST)
h.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax); h.set_state(st)
rc = h.run_days(np.ones(20, np.int32), f)
s = h.get_state()
pickle.dump((rc, {n: getattr(s, n) for n in s.names()}, h.get_annual(1)), open(sys.argv[1], "wb"))
'''
outs = []
for v in sys.argv[1:]:
    env = dict(os.environ, H9GPU_LIB=os.path.abspath(f"variants/libh9gpu_{v}.so"))
    out = f"/tmp/cmp_{v}.pkl"
    subprocess.run([sys.executable, "-c", code, out], check=True, env=env)
    outs.append(pickle.load(open(out, "rb")))
a, b = outs
print("rc", a[0], b[0])
print("state identical:", all(np.array_equal(a[1][n], b[1][n], equal_nan=True) for n in a[1]))
print("annual identical:", all(np.array_equal(a[2][k], b[2][k], equal_nan=True) for k in a[2]))
