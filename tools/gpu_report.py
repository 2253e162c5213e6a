"""Development aid: error statistics GPU vs oracle, written to gpurun_out/ so the
tolerances in tests/ can be calibrated from measured numbers."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py  # noqa: E402
from hybrid9_b200 import H9, MATH_EXACT, MATH_FAST, synth  # noqa: E402
from hybrid9_b200.state import init_state  # noqa: E402


def stats(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    d = np.abs(a - b)
    rel = d / np.maximum(np.abs(b), 1e-30)
    return {"max_abs": float(np.nanmax(d)), "max_rel": float(np.nanmax(rel[np.abs(b) > 1e-6])) if (np.abs(b) > 1e-6).any() else 0.0,
            "p999_abs": float(np.nanquantile(d, 0.999)), "mean_abs": float(np.nanmean(d))}


def compare(w, st0, forcing, nd, nis, mode, label, out, kind="f32"):
    yi = np.ones(nd, np.int32)
    o = oracle_py.Oracle(kind)
    o.configure(w.nx, w.ny, nis, synth.ZI_DRIVER, nyr=1)
    o.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax)
    o.set_state(st0)
    o.set_options(loop_order=1)
    t = time.time()
    orc = o.run_days(yi, forcing)
    t_or = time.time() - t
    ref = o.get_state()
    oann = o.get_annual(1)
    h = H9(0)
    h.configure(w.nx, w.ny, nis, synth.ZI_DRIVER, nyr=1)
    h.set_math(mode)
    h.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax)
    h.set_state(st0)
    t = time.time()
    rc = h.run_days(yi, forcing)
    t_gpu = time.time() - t
    got = h.get_state()
    gann = h.get_annual(1)
    land = w.land
    ofl = o.get_fault()
    res = {"oracle_rc": orc, "gpu_rc": rc, "t_oracle": t_or, "t_gpu": t_gpu,
           "oracle_nfault": ofl["n_faulted"], "gpu_nfault": h.get_fault().n_faulted}
    for n in ("h2osoi_liq", "zwt", "wa", "lai", "lai_litter", "plant_mass", "plant_foliage_mass",
              "rootr_col", "smp"):
        res[n] = stats(getattr(got, n)[land], getattr(ref, n)[land])
    for n in ("npp", "plant_mass", "rnf", "theta_total", "theta"):
        res["axy_" + n] = stats(gann[n][land], oann[n][land])
    out[label] = res
    h.close()
    o.close()


def main():
    out = {}
    w = synth.make_world(nx=144, ny=72, seed=5)
    nis = 48
    f = synth.make_forcing(w, 30, seed=3)
    st_init = init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER)
    st_rand = synth.randomize_state(w, st_init, seed=11)
    f1 = {k: np.ascontiguousarray(v[:1]) for k, v in f.items()}
    for mode, mname in ((MATH_EXACT, "exact"), (MATH_FAST, "fast")):
        compare(w, st_rand, f1, 1, 1, mode, f"{mname}_rand_1step", out)
        compare(w, st_rand, f1, 1, nis, mode, f"{mname}_rand_1day", out)
        compare(w, st_init, f, 30, nis, mode, f"{mname}_init_30day", out)
        compare(w, st_rand, f, 30, nis, mode, f"{mname}_rand_30day", out)
    # rounding-noise floor: float oracle vs double oracle
    for lab, st, nd in (("noise_init_30day", st_init, 30), ("noise_rand_30day", st_rand, 30)):
        yi = np.ones(nd, np.int32)
        res = {}
        sts = {}
        for kind in ("f32", "f64"):
            o = oracle_py.Oracle(kind)
            o.configure(w.nx, w.ny, nis, synth.ZI_DRIVER, nyr=1)
            o.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax)
            o.set_state(st)
            o.set_options(loop_order=1)
            o.run_days(yi, f)
            sts[kind] = o.get_state()
            o.close()
        for n in ("h2osoi_liq", "zwt", "wa", "lai", "plant_mass", "smp"):
            res[n] = stats(getattr(sts["f32"], n)[w.land], getattr(sts["f64"], n)[w.land])
        out[lab] = res
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_report.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    for k, v in out.items():
        print(k, {a: (b if not isinstance(b, dict) else (b["max_abs"], b["max_rel"])) for a, b in v.items()})


if __name__ == "__main__":
    main()
