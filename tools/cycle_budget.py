"""clock() budget of one sub-step of the thread-per-cell fast kernel in the lone-warp regime
(one latitude band of an 8-way split of the 0.5 deg grid: at most one warp per scheduler).

Builds variants/libh9gpu_cycles.so with -DH9_CYCLE_BUDGET (H9_TICK fences in
h9_physics_fast.cuh; not in the product library), steps `--days` days of band `--band` and
prints the average cycles per sub-step of each segment for the warp that holds `--cell`.
The ticks are scheduling fences, so the segments cannot overlap: their sum is an upper bound
of the unfenced sub-step, which is printed beside it (same launch shape, product library).

usage (GPU box):  python tools/cycle_budget.py --band 0 --cell 0 [--days 60]"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SEG = ["block A head: theta, w0, fsat, beta, energy balance, infiltration, aquifer-layer equilibrium (:141-508,576-590)",
       "block A: eight layers: equilibrium profile, hk/smp, fluxes, sweep rows 1-7 (:517-573,598-735,806-826)",
       "block A end: aquifer node and interface, sweep rows 8-9 (:645-650,737-799)",
       "block B: recharge, drainage, baseflow, back substitution, clamp, repair trigger, balance (:828-1283)",
       "the rarely taken cascade / dryness repair (:1131-1211)",
       "-",
       "fault bookkeeping, loop"]


def run(lib, band, days, cell, block, spin=3):
    env = dict(os.environ, H9GPU_LIB=lib, H9_BUDGET_CELL=str(cell))
    code = f"""
import sys, ctypes as C, json, numpy as np
sys.path.insert(0, {ROOT!r})
from hybrid9_b200 import H9, MATH_FAST, synth
from hybrid9_b200.state import init_state
from hybrid9_b200 import distributed as h9d
import torch
w = synth.make_world()
w = h9d.shard_world(w, {band}, 8)[0]
nd = {days}
f = synth.make_forcing(w, nd, seed=9)
h = H9(0); h.configure(w.nx, w.ny, 48, synth.ZI_DRIVER, nyr=1); h.set_math(MATH_FAST); h.set_tuning(0, {block})
h.set_soil(w.soil_tex, w.theta_s, w.hksat, w.bsw, w.psi_s, w.fmax)
h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), with_smp=False)
p, ds, ps = h.pack_forcing(f, nd)
yi = np.ones(nd, np.int32)
for _ in range({spin}): h.run_days_device(yi, p, ds, ps)   # reach the regime of the bench (water tables settle)
h.reset_counters()
h.run_days_device(yi, p, ds, ps)
ms = h.counters()["step_kernel_ms"]
st = h.get_state()
zw = st.zwt[w.land]
warp = {cell} // 32
shallow_in_warp = int((zw[warp*32:warp*32+32] <= 2.296).sum())
out = dict(ms=ms, variant=h.kernel_variant(), nc=h.num_land, shallow_in_warp=shallow_in_warp,
           shallow_share=float((zw <= 2.296).mean()))
try:
    buf = (C.c_ulonglong * 14)()
    rc = h.lib.h9_debug_cycle_budget(buf)
    out["budget"] = list(buf) if rc == 0 else None
    cy, rp = (C.c_uint * 4096)(), (C.c_uint * 4096)()
    h.lib.h9_debug_warp_cycles(cy, rp)
    sm = (C.c_uint * 4096)()
    h.lib.h9_debug_warp_smid(sm)
    out["warp_smid"] = list(sm)[:(h.num_land + 31) // 32]
    nw = (h.num_land + 31) // 32
    ge, sl = (C.c_uint * 4096)(), (C.c_uint * 4096)()
    h.lib.h9_debug_warp_general(ge, sl)
    out["warp_general"] = list(ge)[:nw]
    out["warp_slow"] = [list(sl)[:nw], list(sl)[1024:1024 + nw], list(sl)[2048:2048 + nw]]
    nw = (h.num_land + 31) // 32
    out["warp_cycles"] = list(cy)[:nw]
    out["warp_repairs"] = list(rp)[:nw]
    out["warp_shallow"] = [int((zw[k*32:k*32+32] <= 2.296).sum()) for k in range(nw)]
    L = w.land
    feats = {{"bsw_max": w.bsw[L].max(axis=1), "bsw_min": w.bsw[L].min(axis=1), "ths_min": w.theta_s[L].min(axis=1),
             "hksat_max": w.hksat[L].max(axis=1), "hksat_min": w.hksat[L].min(axis=1), "psi_min": w.psi_s[L].min(axis=1),
             "zwt": zw, "lai": st.lai[L], "h2o_top": st.h2osoi_liq[L][:, 0], "h2o_min": st.h2osoi_liq[L].min(axis=1),
             "smp_min": st.smp[L].min(axis=1), "wa": st.wa[L], "pr_mean": f["pr"][:, L].mean(axis=0),
             "rsds_mean": f["rsds"][:, L].mean(axis=0), "tas_mean": f["tas"][:, L].mean(axis=0),
             "row": np.nonzero(L)[0].astype(np.float32)}}
    out["feat"] = {{k: [[float(v[j*32:j*32+32].min()), float(v[j*32:j*32+32].max())] for j in range(nw)] for k, v in feats.items()}}
except AttributeError:
    out["budget"] = None
print("RESULT " + json.dumps(out))
"""
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    for line in r.stdout.splitlines():
        if line.startswith("RESULT "):
            import json
            return json.loads(line[7:])
    raise SystemExit(r.stdout[-2000:] + r.stderr[-4000:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--band", type=int, default=0)
    ap.add_argument("--cell", type=int, default=0)
    ap.add_argument("--days", type=int, default=60)
    ap.add_argument("--block", type=int, default=64)
    ap.add_argument("--mhz", type=float, default=1965.0)
    ap.add_argument("--spin", type=int, default=3, help="launches of --days days before the measured one")
    ap.add_argument("--coarse", action="store_true", help="leave block A unfenced inside (-DH9_CYCLE_BUDGET=2)")
    a = ap.parse_args()
    name = "cycles2" if a.coarse else "cycles"
    lib = os.path.join(ROOT, "variants", f"libh9gpu_{name}.so")
    if not os.path.exists(lib):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "build_variant.py"), name,
                        "-DH9_CYCLE_BUDGET=2" if a.coarse else "-DH9_CYCLE_BUDGET=1"], check=True)
    if a.coarse:
        SEG[2] = "block A as one piece (segments 0-2 free to overlap)"
        SEG[0] = SEG[1] = "-"
    fenced = run(lib, a.band, a.days, a.cell, a.block, a.spin)
    plain = run(os.path.join(ROOT, "hybrid9_b200", "libh9gpu.so"), a.band, a.days, a.cell, a.block, a.spin)
    nsub = a.days * 48
    print(f"band {a.band} of 8 ({fenced['nc']} cells, {100 * fenced['shallow_share']:.1f} % with the water table "
          f"inside the soil column), kernel {fenced['variant']}, warp of cell {a.cell}: "
          f"{fenced['shallow_in_warp']} of its 32 cells shallow; {a.days} days x 48 sub-steps")
    b = fenced["budget"]
    tot = 0.0
    for k, name in enumerate(SEG):
        if name == "-":
            continue
        cyc = b[k] / nsub
        tot += cyc
        print(f"  {cyc:8.1f} cycles  {name}")
    print(f"  {tot:8.1f} cycles  sum of the fenced segments")
    print(f"  {b[12] / nsub:8.1f} cycles  per sub-step, whole kernel of that thread, fenced build "
          f"(includes GROW and the daily bookkeeping, 1/48 each)")
    print(f"  {fenced['ms'] * 1e-3 * a.mhz * 1e6 / nsub:8.1f} cycles  per sub-step from the launch time of the fenced build "
          f"({fenced['ms']:.3f} ms: the slowest warp of the shard)")
    print(f"  {plain['ms'] * 1e-3 * a.mhz * 1e6 / nsub:8.1f} cycles  per sub-step from the launch time of the product library "
          f"({plain['ms']:.3f} ms), segments free to overlap")
    if fenced.get("warp_cycles"):
        import numpy as np
        wc = np.array(fenced["warp_cycles"], float) / nsub
        rp = np.array(fenced["warp_repairs"], float) / nsub
        sh = np.array(fenced["warp_shallow"])
        print(f"  per warp of the fenced build ({len(wc)} warps), cycles per sub-step: min {wc.min():.0f}, median "
              f"{np.median(wc):.0f}, p90 {np.quantile(wc, 0.9):.0f}, max {wc.max():.0f}")
        order = np.argsort(-wc)[:8]
        smid = np.array(fenced["warp_smid"]) >> 8
        wid = np.array(fenced["warp_smid"]) & 0xff
        per_sm = np.bincount(smid, minlength=148)
        sched = smid * 4 + (wid % 4)
        per_sched = np.bincount(sched)
        ge = np.array(fenced["warp_general"], float) / nsub
        sl = np.array(fenced["warp_slow"], float) / nsub
        for k in order:
            print(f"    warp {k:4d}: {wc[k]:7.0f} cycles, {sh[k]:2d} shallow cells at the end; general step in "
                  f"{100 * ge[k]:5.1f} % of its sub-steps; most affected lane: recharge loop went on in "
                  f"{100 * sl[0][k]:5.1f} %, baseflow loop in {100 * sl[1][k]:5.1f} %, bottom-layer search in "
                  f"{100 * sl[2][k]:5.1f} %, cascade/repair branch in {100 * rp[k]:5.1f} %")
        gen = ge > 0.5
        if gen.any() and (~gen).any():
            print(f"  warps mostly on the all-deep step: {int((~gen).sum())}, mean {wc[~gen].mean():.0f} cycles; mostly on the "
                  f"general step: {int(gen.sum())}, mean {wc[gen].mean():.0f}, max {wc[gen].max():.0f}")
        ft = fenced.get("feat") or {}
        for name, mm in ft.items():
            mm = np.array(mm)
            cmin, cmax = np.corrcoef(wc, mm[:, 0])[0, 1], np.corrcoef(wc, mm[:, 1])[0, 1]
            print(f"    corr(cycles, warp-min {name}) = {cmin:+.2f}, corr(cycles, warp-max {name}) = {cmax:+.2f}; "
                  f"slowest warp: [{mm[order[0], 0]:.4g}, {mm[order[0], 1]:.4g}], all warps: [{mm[:, 0].min():.4g}, {mm[:, 1].max():.4g}]")
        print(f"  SMs in use: {int((per_sm > 0).sum())}; warps per SM: max {per_sm.max()}; schedulers holding 2+ warps: "
              f"{int((per_sched > 1).sum())}; mean cycles of warps alone on their scheduler "
              f"{wc[per_sched[sched] == 1].mean():.0f}, of warps sharing one {wc[per_sched[sched] > 1].mean() if (per_sched[sched] > 1).any() else float('nan'):.0f}")
        print(f"  warps that never take the repair branch: {int((rp == 0).sum())}; correlation(cycles, repair share) = "
              f"{np.corrcoef(wc, rp)[0, 1]:.2f}")


if __name__ == "__main__":
    main()
