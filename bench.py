#!/usr/bin/env python
"""bench.py -- land-cell-timesteps/s of the HYDROLOGY+GROW path (BASELINE.json metric).

One "step" = one simulated year (365 days x NISURF=48 sub-steps) over every land
cell of the rank's block: forcing derivation, 48 x HYDROLOGY, GROW, annual
accumulators -- the loop nest HYBRID9.f90:120-295.  Workload at N=1: the 0.5 deg
global land mask (67,420 cells), configs[2] of BASELINE.json, on synthetic
PGF-shaped forcing (hybrid9_b200/synth.py).

  value  device-resident: compact forcing already in HBM, one fused launch per
         year (h9_run_days_device); CUDA events on the launching stream.
  e2e    through h9_run_days with the seven (lon_c,lat_c,ndays) HOST arrays as
         READ_PGF leaves them (pinned), host->device copies and the device-side
         pack inside the timed region, plus the device->host read of the year's
         annual means (h9_get_annual) and the fault word.
  roofline  352 algorithmic bytes per land-cell-timestep over the
         fused kernel's CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's own Fortran, translated statement for statement to C++ by
         oracle/f2cpp.py and compiled -O3 (oracle/_ref; kind "reference"; the image has no
         Fortran compiler), on all host cores, bounded sample; the hand-written port beside it.

  roofline.pipe  the binding roofline: issue slots and MUFU of the fused kernel, from the
         instruction counts of the committed ncu capture (profiles/pipe.json).

`--impl reference` times that CPU implementation alone (rank 0), same metric/config; both
CPU figures time the SAME fixed sample (cpu_sample) and report the median of the repeats.
Multi-GPU (torchrun, one process per GPU): ONE 0.5 deg grid sharded in contiguous latitude
bands (`--scaling strong`, the default: BASELINE.json's metric is one global grid on
1/2/4/8 GPUs); the replicated-grid number (every rank a full block, `--scaling weak`) is
measured in the same run and carried as the `weak` sub-record.  The only collectives are the
per-year annual-mean all-gather and the FP64 budget all-reduce, issued by the library itself
(h9_annual_collective: NCCL on the ctx's stream) inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_CELL_STEP = 352.0
METRIC = "land_cell_timesteps_per_s"
UNIT = "cell-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling sub-record at N>1")
    ap.add_argument("--grid", default="0.5", choices=["0.5", "0.25", "regional", "tiny", "band8"])
    ap.add_argument("--days", type=int, default=365)
    ap.add_argument("--nisurf", type=int, default=48)
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--tile-days", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-repeats", type=int, default=5)
    return ap.parse_args()


def grid_spec(name):
    if name == "0.5":
        return dict(nx=720, ny=360, n_land=67420, label="0.5deg global land mask")
    if name == "0.25":
        return dict(nx=1440, ny=720, n_land=269680, label="0.25deg global land mask")
    if name == "regional":
        # SURVEY.md 8d names a 100x74 window at (lon_s=340, lat_s=37) ("Europe"); the synthetic
        # continents are not Earth's, so the window is placed where it holds 2,500 land cells
        return dict(nx=720, ny=360, n_land=67420, label="regional 100x74 window of the 0.5deg mask "
                    "(2,500 land cells)", window=synth_regional_window())
    if name == "band8":
        nb = int(os.environ.get("H9_BENCH_NBANDS", "8"))
        k = int(os.environ.get("H9_BENCH_BAND", "0"))
        return dict(nx=720, ny=360, n_land=67420, band=(k, nb),
                    label=f"band {k} of {nb} latitude bands of the 0.5deg mask (one GPU's share of a "
                          f"{nb}-GPU strong-scaling run)")
    return dict(nx=72, ny=36, n_land=674, label="tiny 5deg test grid")


def synth_regional_window():
    from hybrid9_b200 import synth
    return synth.REGIONAL_WINDOW


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_engines(world, forcing, ncell, nisurf, cores):
    """The CPU implementations of the path that exist on this box, fastest build of each:
    "reference" = the reference's own HYDROLOGY.f90 / GROW.f90 / loop nest, translated statement
    for statement by oracle/f2cpp.py and compiled -O3 (oracle/_ref/libh9ref_o3.so), one
    independent instance per host thread like its MPI ranks; "port" = the hand-written C++
    restatement (oracle/libh9oracle_o3.so).  Each entry: run(ndays) -> seconds."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py
    import ref_py
    from hybrid9_b200 import synth
    eng = {}
    if ref_py.available("o3"):
        pool = ref_py.RefRanks(world, forcing, ncell, nisurf, cores, kind="o3")
        eng["reference"] = (pool.run_days,
                            "the reference's own HYDROLOGY.f90/GROW.f90/HYBRID9.f90:120-295 translated "
                            "statement for statement to C++ (oracle/f2cpp.py; no Fortran compiler in the "
                            f"image), g++ -O3 -march=x86-64-v3, {cores} independent blocks like MPI ranks")
    cw = synth.compact_world(world, ncell)

    def run_port(nd):
        cf = synth.compact_forcing(world, forcing, ncell, ndays=nd)
        o = oracle_py.Oracle("o3")
        o.configure(cw.nx, cw.ny, nisurf, synth.ZI_DRIVER, nyr=1)
        o.set_soil(cw.soil_tex, cw.theta_s, cw.hksat, cw.bsw, cw.psi_s, cw.fmax)
        o.init_state()
        o.set_options(loop_order=0, smp_leak=0, nthreads=cores)  # the reference's cell-outer order
        t = time.perf_counter()
        o.run_days(np.ones(nd, np.int32), cf)
        dt = time.perf_counter() - t
        o.close()
        return dt
    eng["port"] = (run_port, "hand-written C++ restatement of HYDROLOGY.f90/GROW.f90 "
                             f"(oracle/h9_oracle.cpp), g++ -O3 -march=x86-64-v3, {cores} threads over cells")
    return eng


CPU_SAMPLE_CELLS_PER_CORE = 1024
CPU_SAMPLE_DAYS = 24


def cpu_sample(args):
    """The ONE bounded sample both CPU legs time: the first 1024 x cores land cells of the
    0.5 deg mask (reference iteration order) x 24 days x NISURF sub-steps, INIT state,
    forcing seed 9.  Fixed cells x days, so `cpu_baseline` of the graft arm and the line of
    `--impl reference` time the same work."""
    from hybrid9_b200 import synth
    g = grid_spec("0.5")
    world = synth.make_world(nx=g["nx"], ny=g["ny"], n_land=g["n_land"], seed=9)
    cores = os.cpu_count() or 1
    ncell = int(min(world.land.sum(), CPU_SAMPLE_CELLS_PER_CORE * cores))
    cw = synth.compact_world(world, ncell)
    forcing = synth.make_forcing(cw, CPU_SAMPLE_DAYS, seed=9)
    return cw, forcing, ncell, cores


def cpu_time_sample(args, repeats, warm, kinds=("reference", "port")):
    """Median-of-repeats rate of each CPU engine on cpu_sample().  Returns {kind: {...}}."""
    cw, forcing, ncell, cores = cpu_sample(args)
    eng = cpu_engines(cw, forcing, ncell, args.nisurf, cores)
    nd = CPU_SAMPLE_DAYS
    units = ncell * nd * args.nisurf
    out = {}
    for kind in kinds:
        if kind not in eng:
            continue
        run, what = eng[kind]
        for _ in range(max(1, warm)):  # first touch of the arrays, thread start-up
            run(nd)
        ts = sorted(run(nd) for _ in range(max(3, repeats)))
        med = ts[len(ts) // 2]
        out[kind] = {"value": units / med, "seconds_median": med, "seconds_min": ts[0],
                     "seconds_max": ts[-1], "spread": (ts[-1] - ts[0]) / med, "repeats": len(ts),
                     "units_per_repeat": units,
                     "sample": f"{ncell} land cells x {nd} days x {args.nisurf} sub-steps (fixed), "
                               f"median of {len(ts)} repeats, cell-outer loop order; {what}"}
    return out, cores


def cpu_baseline(args):
    """Time the reference's CPU implementation on all host cores over the fixed sample.
    `value` is the translated reference when it is present (kind "reference"); the
    hand-written port is timed beside it."""
    out, cores = cpu_time_sample(args, args.cpu_repeats, 1)
    kind = "reference" if "reference" in out else "port"
    res = {"value": out[kind]["value"], "unit": UNIT, "cores": cores, "kind": kind,
           "sample": out[kind]["sample"], "seconds": out[kind]["seconds_median"],
           "spread": out[kind]["spread"], "repeats": out[kind]["repeats"]}
    if kind == "reference" and "port" in out:
        res["port_value"] = out["port"]["value"]
        res["port_sample"] = out["port"]["sample"]
    return res


def build_world(args):
    from hybrid9_b200 import synth
    g = grid_spec(args.grid)
    w = synth.make_world(nx=g["nx"], ny=g["ny"], n_land=g["n_land"], seed=9)
    if "window" in g:
        w = w.window(*g["window"])
    if "band" in g:
        from hybrid9_b200 import distributed as h9d
        w = h9d.shard_world(w, g["band"][0], g["band"][1])[0]
    return w, g["label"]


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path alone (the
    translated Fortran when oracle/_ref is present, else the hand-written port), rank 0 only,
    all host threads; each step is one pass over the fixed sample of cpu_sample()."""
    if rank != 0:
        return
    label = grid_spec(args.grid)["label"]
    kinds = ("reference",)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_py
    if not ref_py.available("o3"):
        kinds = ("port",)
    out, cores = cpu_time_sample(args, args.steps, max(1, args.warmup), kinds=kinds)
    kind = kinds[0]
    r = out[kind]
    value, sample = r["value"], r["sample"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["repeats"], "warmup": max(1, args.warmup), "ms_per_step": 1e3 * r["seconds_median"],
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{label}, {args.days} d x {args.nisurf} sub-steps "
                               f"(bounded CPU sample: {sample})"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "spread": r["spread"], "repeats": r["repeats"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else that libraries
    write to fd 1 (NCCL's version banner under NCCL_DEBUG=VERSION ...) was sent to stderr."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # C-level and Python-level stdout of this process now go to stderr
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the HYDROLOGY/GROW path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    multi = world_size > 1
    if multi:
        dist.init_process_group("nccl", device_id=dev)

    res = measure(args, args.scaling, rank, local_rank, world_size, dev, e2e=not args.no_e2e,
                  steps=args.steps, sample_clocks=True)
    weak = None
    if multi and args.scaling == "strong" and not args.no_weak:
        # the replicated-grid number of the same run: every rank steps a full block
        w = measure(args, "weak", rank, local_rank, world_size, dev, e2e=False,
                    steps=min(args.steps, 5), sample_clocks=False)
        weak = {"value": w["value"], "unit": UNIT, "ms_per_step": w["ms_per_step"], "steps": w["steps"],
                "land_cells_per_gpu": w["n_land_all"], "scaling": "weak",
                "note": "every rank steps its own full 0.5deg block (replicas); secondary"}

    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline(args)
        except Exception as ex:  # the checker is test infrastructure; never fail the bench on it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"failed: {ex}"}

    if rank == 0:
        nd, nis = args.days, args.nisurf
        n_land_all = res["n_land_all"]
        sharded = multi and args.scaling == "strong"
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": (f"{res['label']} sharded in contiguous latitude bands over "
                                    f"{world_size} GPUs: {sum(n_land_all)} land cells x {nd} days x "
                                    f"{nis} sub-steps per step") if sharded else
                                   (f"{res['label']}: {n_land_all[0]} land cells x {nd} days x {nis} "
                                    f"sub-steps per step" + (" per GPU" if multi else "")),
                       "land_cells_per_gpu": n_land_all, "days_per_step": nd, "nisurf": nis,
                       "math": args.math, "simulated_years": res["years"],
                       "share_cells_water_table_in_soil_column_at_end": round(res["shallow"], 4),
                       "l2": "forcing stream (689 MB/step at 0.5deg) exceeds L2; state is "
                             "register/L2 resident by design",
                       "kernel_variant": res["variant"],
                       "parallelism": f"dp{world_size} latitude bands, no data-path collective; per "
                                      "simulated year one FP64 budget all-reduce and one all-gather "
                                      "of the annual means (h9_annual_collective, NCCL on the "
                                      "library's stream)" if multi else "single GPU"},
            "roofline": res["roofline"], "cpu_baseline": cpu, "e2e": res["e2e"],
            "gpu_launches": res["launches"], "clocks": res["clocks"], "wall_s_timed": res["wall"],
        }
        if weak is not None:
            line["weak"] = weak
        emit(line)
    if multi:
        dist.destroy_process_group()


def pipe_roofline(variant, units_rank, kern_ms, clocks):
    """The binding roofline of the fused kernel: issue slots (one warp instruction per
    scheduler per cycle) and the MUFU unit (8 cycles per warp instruction per scheduler),
    from the executed-instruction counts of the committed ncu capture (profiles/pipe.json),
    at the SM clock sampled during the run."""
    pf = os.path.join(ROOT, "profiles", "pipe.json")
    if not os.path.exists(pf):
        return None
    try:
        with open(pf) as f:
            pj = json.load(f)
        # every launch shape of one build executes the same instructions: <BLOCK,1> is the
        # all-registers build, <BLOCK,4|8|16> the 128-register one
        import re
        fam = re.sub(r"<(MathExact,)?\d+,1>", r"<\g<1>64,1>", variant or "")
        fam = re.sub(r"<(MathExact,)?\d+,(4|8|16)>", r"<\g<1>64,8>", fam)
        k = pj["kernels"].get(fam)
        if k is None and "MathExact" in fam:
            return None  # no capture of that build: no floor rather than another kernel's
        k = k or pj["kernels"][pj["default"]]
        mhz = (clocks or {}).get("sm_mhz") or pj.get("sm_mhz", 1965.0)
        sched = pj.get("schedulers", 148 * 4)
        wsub = units_rank / float(k["cell_steps_per_warp_substep"])  # warp-sub-steps per launch
        issue_ms = wsub * k["warp_inst_per_substep"] / (sched * mhz * 1e3)
        mufu_ms = wsub * k["mufu_per_substep"] * 8.0 / (sched * mhz * 1e3)
        floor_ms = max(issue_ms, mufu_ms)
        return {"bound": "issue" if issue_ms >= mufu_ms else "mufu",
                "warp_inst_per_substep": k["warp_inst_per_substep"],
                "mufu_per_substep": k["mufu_per_substep"],
                "cell_steps_per_warp_substep": k["cell_steps_per_warp_substep"],
                "issue_floor_ms": issue_ms, "mufu_floor_ms": mufu_ms, "floor_ms": floor_ms,
                "frac": floor_ms / kern_ms if kern_ms > 0 else None, "sm_mhz": mhz,
                "schedulers": sched, "source": k.get("source")}
    except Exception as ex:
        return {"error": str(ex)}


def measure(args, scaling, rank, local_rank, world_size, dev, e2e, steps, sample_clocks):
    """One workload (the whole grid on one GPU, a latitude band of it, or a replica) through
    the device-resident and the end-to-end entry; returns the numbers of this rank's line."""
    import torch
    import torch.distributed as dist
    from hybrid9_b200 import H9, MATH_EXACT, MATH_FAST, synth
    from hybrid9_b200.host import pinned_empty
    from hybrid9_b200.state import init_state
    from hybrid9_b200 import distributed as h9d

    multi = world_size > 1
    world, label = build_world(args)
    if multi and scaling == "strong":
        world, _, _, _ = h9d.shard_world(world, rank, world_size)
    nd, nis = args.days, args.nisurf
    nc = int(world.land.sum())

    # host forcing exactly as READ_PGF leaves it, in pinned memory (pageable for the
    # device-resident-only weak sub-record)
    names = ("tas", "rlds", "rsds", "huss", "ps", "pr", "rhs")
    if e2e:
        forcing = {k: pinned_empty((nd, world.ny, world.nx)) for k in names}
    else:
        forcing = {k: np.empty((nd, world.ny, world.nx), np.float32) for k in names}
    synth.make_forcing(world, nd, seed=9 + rank, out=forcing)

    h = H9(local_rank)
    nyr = 2
    h.configure(world.nx, world.ny, nis, synth.ZI_DRIVER, nyr=nyr)
    h.set_math(MATH_EXACT if args.math == "exact" else MATH_FAST)
    h.set_tuning(args.tile_days, args.block)
    h.set_soil(world.soil_tex, world.theta_s, world.hksat, world.bsw, world.psi_s, world.fmax)
    assert h.num_land == nc
    h.set_state(init_state(world.soil_tex, world.theta_s, synth.ZI_DRIVER), with_smp=False)
    n_land_all = [nc]
    if multi:
        # the library's own NCCL communicator; torch.distributed only carries the 128-byte id
        n_land_all = [int(x) for x in h9d.comm_init_from_torch(h, device=dev)]
    d_forc, day_stride, plane_stride = h.pack_forcing(forcing, nd)
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    year = [0]

    def collective(iy):
        if multi:
            h.annual_collective(iy)  # stream-ordered, no host sync
            if year[0] <= 1:  # sanity of the gathered view, outside the steady state
                parts, budget = h9d.fetch_gathered(h, iy, n_land_all)
                assert [int(p.shape[1]) for p in parts] == n_land_all
                assert int(round(float(budget[5]))) == sum(n_land_all) and float(budget[7]) == 0.0

    def step_device():
        year[0] += 1
        iy = (year[0] - 1) % nyr + 1
        rc = h.run_days_device(np.full(nd, iy, np.int32), d_forc, day_stride, plane_stride)
        if rc:
            raise SystemExit(f"physics fault {rc}: {h.get_fault()}")
        collective(iy)

    def step_e2e():
        year[0] += 1
        iy = (year[0] - 1) % nyr + 1
        rc = h.run_days(np.full(nd, iy, np.int32), forcing)
        if rc:
            raise SystemExit(f"physics fault {rc}: {h.get_fault()}")
        ann = h.get_annual(iy)
        collective(iy)
        return ann

    def timed(fn, nsteps, warmup, clocks=False):
        for _ in range(warmup):
            fn()
        h.synchronize()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        h.reset_counters()
        cs = ClockSampler(local_rank) if clocks else None
        if cs:
            cs.start()
            time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(nsteps):
            fn()
        if multi:
            h.collective_fence()  # the last year's collectives (own stream) end inside the timed region
        e1.record(stream)
        h.synchronize()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if multi:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        ck = cs.stop() if cs else None
        cnt = h.counters()
        if multi:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, cnt, ck

    units_rank = nc * nd * nis
    units_all = sum(n_land_all) * nd * nis

    ms, wall, cnt, clocks = timed(step_device, steps, max(args.warmup, 3), clocks=sample_clocks)
    ms_per_step = ms / steps
    value = units_all / (ms_per_step * 1e-3)
    kern_ms = cnt["step_kernel_ms"] / max(1, steps)  # fused kernel, CUDA events, this rank
    peak, peak_src = peaks()
    achieved = units_rank * BYTES_PER_CELL_STEP / (kern_ms * 1e-3) / 1e9
    variant = h.kernel_variant()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "kernel": variant,
                "kernel_ms_per_launch": kern_ms, "launches_per_step": cnt["launches"] / steps,
                "algorithmic_bytes_per_unit": BYTES_PER_CELL_STEP, "peak_source": peak_src,
                "note": "algorithmic roofline of a per-sub-step operator (this rank's shard "
                        "against one GPU's peak); the fused kernel keeps state in registers, real "
                        "DRAM traffic is the forcing stream (see profiles/), the binding limiters "
                        "are issue slots and the MUFU pipe: see `pipe`"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            # measured dram bytes per land-cell-timestep from the committed ncu capture
            roofline["traffic"] = float(tj["dram_bytes_per_cell_step"]) * units_rank
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass
    roofline["pipe"] = pipe_roofline(variant, units_rank, kern_ms, clocks)
    launches = cnt["launches"]

    e2e_rec = None
    if e2e:
        ems, ewall, ecnt, _ = timed(step_e2e, steps, max(args.warmup, 3))
        e_ms = max(ems, ewall * 1e3) / steps
        e2e_rec = {"value": units_all / (e_ms * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": ecnt["h2d_bytes"] // steps,
                   "d2h_bytes_per_step": ecnt["d2h_bytes"] // steps,
                   "ms_per_step": e_ms, "gpu_launches": ecnt["launches"],
                   "api": "h9_run_days (pinned host forcing as READ_PGF leaves it) + h9_get_annual"
                          + (" + h9_annual_collective" if multi else "")}

    # regime of the final state (outside the timed region): share of cells whose water table
    # is inside the soil column (jwt < 8), which take the branchier Drainage path
    st_end = h.get_state()
    shallow = float((st_end.zwt[world.land] <= synth.ZI_DRIVER[8] / 1000.0).mean()) if nc else 0.0
    h.close()
    return {"value": value, "ms_per_step": ms_per_step, "steps": steps, "n_land_all": n_land_all,
            "roofline": roofline, "e2e": e2e_rec, "launches": launches, "clocks": clocks,
            "wall": wall, "years": year[0], "shallow": shallow, "label": label, "variant": variant}


if __name__ == "__main__":
    main()
