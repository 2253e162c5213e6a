"""Calibration data for tests/test_gpu_fullsize_oracle.py (GPU box): quantiles of the
differences fast (thread-per-cell, two lanes) vs exact on the GPU, and of the FP32 noise floor
(oracle float vs double), for (a) the whole 0.5 deg grid after 10 days, (b) year 30 of a spin-up
of a 2.7k-cell block from randomised states.  Writes gpurun_out/equilibrium_report.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import THREAD_PER_CELL, TWO_LANES, make_gpu, make_oracle  # noqa: E402
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth  # noqa: E402
from hybrid9_b200.state import init_state  # noqa: E402

Q = (0.5, 0.9, 0.99, 0.999, 1.0)


def quant(a, b, floor):
    a, b = a.astype(np.float64), b.astype(np.float64)
    r = np.abs(a - b) / np.maximum(np.abs(b), floor)
    return {"rel_q": [float(np.quantile(r, q)) for q in Q], "abs_q": [float(np.quantile(np.abs(a - b), q)) for q in Q],
            "mean_a": float(a.mean()), "mean_b": float(b.mean()),
            "bias_rel": float(abs(a.mean() - b.mean()) / max(abs(b.mean()), 1e-30))}


def compare(tag, sa, aa, sb, ab, land, out):
    res = {}
    for k, floor in (("rnf", 1e-6), ("theta", 1e-3), ("plant_mass", 1e-3), ("theta_total", 1e-3), ("npp", 1e-3)):
        res["axy_" + k] = quant(aa[k][land], ab[k][land], floor)
    for n, floor in (("h2osoi_liq", 1e-3), ("zwt", 1e-3), ("wa", 1e-3), ("lai", 1e-6), ("plant_mass", 1e-3)):
        res[n] = quant(getattr(sa, n)[land], getattr(sb, n)[land], floor)
    out[tag] = res


def main():
    out = {"quantiles": list(Q)}
    nth = os.cpu_count() or 1
    # (b) 30-year spin-up
    w = synth.make_world(nx=144, ny=72, seed=5)
    f = synth.make_forcing(w, 365, seed=3)
    st = synth.randomize_state(w, init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), seed=11)
    land = w.land
    runs = {}
    for name, mode, block in (("exact", MATH_EXACT, 0), ("thread", MATH_FAST, THREAD_PER_CELL), ("pair", MATH_FAST, TWO_LANES)):
        h = make_gpu(w, mode=mode, nyr=2, block=block)
        h.set_state(st)
        p, ds, ps = h.pack_forcing(f, 365)
        t = time.time()
        rc = 0
        for yr in range(30):
            rc |= h.run_days_device(np.full(365, yr % 2 + 1, np.int32), p, ds, ps)
        runs[name] = (h.get_state(), h.get_annual(2), rc, h.get_fault().n_faulted, time.time() - t)
        h.close()
    out["spinup_rc"] = {k: [v[2], v[3], v[4]] for k, v in runs.items()}
    for kind in ("f32", "f64"):
        o = make_oracle(w, kind=kind, nyr=2, nthreads=nth)
        o.set_state(st)
        t = time.time()
        rc = 0
        for yr in range(30):
            rc |= o.run_days(np.full(365, yr % 2 + 1, np.int32), f)
        runs["oracle_" + kind] = (o.get_state(), o.get_annual(2), rc, o.get_fault()["n_faulted"], time.time() - t)
        o.close()
    out["oracle_rc"] = {k: [v[2], v[3], v[4]] for k, v in runs.items() if k.startswith("oracle")}
    compare("spinup30_thread_vs_exact", runs["thread"][0], runs["thread"][1], runs["exact"][0], runs["exact"][1], land, out)
    compare("spinup30_pair_vs_exact", runs["pair"][0], runs["pair"][1], runs["exact"][0], runs["exact"][1], land, out)
    compare("spinup30_exact_vs_oracle32", runs["exact"][0], runs["exact"][1], runs["oracle_f32"][0], runs["oracle_f32"][1], land, out)
    compare("spinup30_noise_oracle32_vs_64", runs["oracle_f32"][0], runs["oracle_f32"][1], runs["oracle_f64"][0], runs["oracle_f64"][1], land, out)
    se = runs["exact"][0]
    out["spinup30_shallow_share_exact"] = float((se.zwt[land] <= 2.296).mean())
    with open(os.path.join(ROOT, "gpurun_out", "equilibrium_report.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    for k, v in out.items():
        if isinstance(v, dict) and "axy_rnf" in v:
            print(k)
            for n, qq in v.items():
                print("   %-16s rel p50 %.2e p99 %.2e p99.9 %.2e max %.2e  bias %.2e" % (n, qq["rel_q"][0], qq["rel_q"][2], qq["rel_q"][3], qq["rel_q"][4], qq["bias_rel"]))
    print(out["spinup_rc"], out["oracle_rc"], out["spinup30_shallow_share_exact"])


if __name__ == "__main__":
    main()
