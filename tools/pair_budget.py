"""clock() budget of the two-lanes-per-cell day kernel per warp (GPU box): cycles per sub-step in
block A and in the tail, separately for the all-deep and the general step, after `--spin` years.
Builds variants/libh9gpu_cycles2.so with -DH9_CYCLE_BUDGET=2 (fences at block A / tail only)."""
import argparse, ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--band", type=int, default=0)
ap.add_argument("--spin", type=int, default=8)
ap.add_argument("--mhz", type=float, default=1965.0)
a = ap.parse_args()
lib = os.path.join(ROOT, "variants", "libh9gpu_cycles2.so")
if not os.path.exists(lib):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "build_variant.py"), "cycles2", "-DH9_CYCLE_BUDGET=2"], check=True)
os.environ["H9GPU_LIB"] = lib
from hybrid9_b200 import H9, MATH_FAST, synth
from hybrid9_b200.state import init_state
from hybrid9_b200 import distributed as h9d
w = h9d.shard_world(synth.make_world(), a.band, 8)[0]
nd = 365
f = synth.make_forcing(w, nd, seed=9)
h = H9(0); h.configure(w.nx, w.ny, 48, synth.ZI_DRIVER, nyr=1); h.set_math(MATH_FAST); h.set_tuning(0, 4000)
h.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax)
h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), with_smp=False)
p, ds, ps = h.pack_forcing(f, nd)
yi = np.ones(nd, np.int32)
for _ in range(a.spin): h.run_days_device(yi, p, ds, ps)
h.reset_counters(); h.run_days_device(yi, p, ds, ps)
ms = h.counters()["step_kernel_ms"]
buf = (C.c_uint * (8 * 4096))()
assert h.lib.h9_debug_pair_budget(buf) == 0
nw = (h.num_land + 15) // 16
b = np.array(buf, dtype=np.float64).reshape(8, 4096)[:, :nw]
nsub = nd * 48
print(f"band {a.band} of 8, {h.num_land} cells, {nw} warps, kernel {h.kernel_variant()}, fenced build {ms:.3f} ms per year = "
      f"{ms * 1e-3 * a.mhz * 1e6 / nsub:.0f} cycles per sub-step for the slowest warp")
nd_, ng = b[4], b[5]
with np.errstate(invalid="ignore", divide="ignore"):
    ad, td, ag, tg = b[0] / nd_, b[1] / nd_, b[2] / ng, b[3] / ng
tot = b[6] / nsub
def stat(x):
    x = x[np.isfinite(x)]
    return "n/a" if x.size == 0 else f"median {np.median(x):.0f}, p90 {np.quantile(x, 0.9):.0f}, max {x.max():.0f}"
print("  all-deep step: block A", stat(ad), "| tail", stat(td))
print("  general step : block A", stat(ag), "| tail", stat(tg))
print("  share of sub-steps on the general step per warp:", stat(100 * ng / nsub), "%")
print("  cycles per sub-step per warp (whole kernel):", stat(tot))
order = np.argsort(-tot)[:5]
for k in order:
    print(f"    warp {k}: {tot[k]:.0f} cycles; general in {100 * ng[k] / nsub:.1f} % of its sub-steps; A/tail general {ag[k]:.0f}/{tg[k]:.0f}, all-deep {ad[k]:.0f}/{td[k]:.0f}")
