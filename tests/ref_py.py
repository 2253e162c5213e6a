"""ctypes wrapper of oracle/_ref/libh9ref.so: the reference's own HYDROLOGY.f90 / GROW.f90 /
loop nest, translated statement for statement by oracle/f2cpp.py and compiled with g++
(oracle/ref_harness.cpp is the C ABI).  TEST INFRASTRUCTURE, like oracle_py: used by tests/,
bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() only.

`RefModel` has the method names and array conventions of `oracle_py.Oracle` / `H9`.
The library is built where /root/reference exists (this container); on the GPU box the
prebuilt file that travelled with the snapshot is used.  `available()` says whether it is
there; tests skip without it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REFERENCE_SRC = "/root/reference/SOURCE"
FORCING = ("tas", "rlds", "rsds", "huss", "ps", "pr", "rhs")
DIAG = ("qflx_surf", "rsub_top", "qflx_rsub_sat", "qflx_infl", "qcharge", "fsat", "beta", "rsc",
        "w0", "w1", "rous", "zwtmm", "desatdT", "gamma", "rho", "Rnets")
_LIBS = {}
_NAMES = {"ref": "libh9ref.so", "o3": "libh9ref_o3.so", "chk": "libh9ref_chk.so",
          "pk": "libh9ref_pk.so"}


def build():
    """(re)translate and compile when the reference sources are present; no-op otherwise."""
    if os.path.isdir(REFERENCE_SRC):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)


def available(kind: str = "ref") -> bool:
    if not os.path.exists(os.path.join(REF_DIR, _NAMES[kind])):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(os.path.join(REF_DIR, _NAMES[kind]))


def load(kind: str = "ref"):
    """kind: ref (strict IEEE, the pin), o3 (CPU-baseline build), chk (bounds-checked), pk (ref with
    the GPU exact mode's portable pow/exp/log instead of glibc's)."""
    if kind in _LIBS:
        return _LIBS[kind]
    if not available(kind):
        raise RuntimeError(f"oracle/_ref/{_NAMES[kind]} is missing and {REFERENCE_SRC} is not "
                           "here to build it from")
    lib = C.CDLL(os.path.join(REF_DIR, _NAMES[kind]))
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    sig = {
        "h9r_create": (C.c_int, [C.POINTER(vp)]),
        "h9r_destroy": (C.c_int, [vp]),
        "h9r_configure": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, fp, C.c_int]),
        "h9r_set_soil": (C.c_int, [vp, ip, fp, fp, fp, fp, fp]),
        "h9r_num_land": (C.c_int64, [vp]),
        "h9r_get_land_index": (C.c_int, [vp, ip]),
        "h9r_init_state": (C.c_int, [vp]),
        "h9r_set_state": (C.c_int, [vp] + [fp] * 10 + [ip, fp]),
        "h9r_get_state": (C.c_int, [vp] + [fp] * 10 + [ip, fp, fp]),
        "h9r_run_days": (C.c_int, [vp, C.c_int, ip] + [fp] * 7 + [C.c_int]),
        "h9r_run_decade": (C.c_int, [vp, C.c_int, C.c_int, C.c_int] + [fp] * 7),
        "h9r_get_annual": (C.c_int, [vp, C.c_int] + [fp] * 6),
        "h9r_get_annual_forcing": (C.c_int, [vp, C.c_int, fp]),
        "h9r_hydrology_step": (C.c_int, [vp] + [fp] * 7 + [fp] * 5 + [ip, fp, C.c_int]),
        "h9r_grow_day": (C.c_int, [vp, fp, fp, fp, fp, C.c_int]),
        "h9r_get_fault": (C.c_int, [vp, C.POINTER(C.c_uint32), ip, ip, ip, ip, fp, ip]),
        "h9r_get_geometry": (C.c_int, [vp, fp, fp, fp]),
        "h9r_time_boy": (C.c_int, [vp, C.c_int]),
        "h9r_regrid_soil_layer": (C.c_int, [C.c_int, C.c_int, C.c_int] + [fp] * 8),
        "h9r_ndiag": (C.c_int, []),
    }
    for n, (res, args) in sig.items():
        fn = getattr(lib, n)
        fn.restype, fn.argtypes = res, args
    assert lib.h9r_ndiag() == len(DIAG)
    _LIBS[kind] = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f(a):
    return np.ascontiguousarray(a, np.float32)


class RefModel:
    def __init__(self, kind: str = "ref", per_cell_smp: bool = True):
        self.lib = load(kind)
        h = C.c_void_p()
        assert self.lib.h9r_create(C.byref(h)) == 0
        self.h = h
        self.per_cell_smp = 1 if per_cell_smp else 0

    def close(self):
        if self.h:
            self.lib.h9r_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, lon_c, lat_c, nisurf, zi, nyr=1):
        self.lon_c, self.lat_c, self.nisurf, self.nyr = lon_c, lat_c, nisurf, nyr
        assert self.lib.h9r_configure(self.h, lon_c, lat_c, nisurf, _p(_f(zi)), nyr) == 0

    def set_soil(self, soil_tex, theta_s, hksat, bsw, psi_s, fmax):
        st = np.ascontiguousarray(soil_tex, np.int32)
        a = [_f(v) for v in (theta_s, hksat, bsw, psi_s, fmax)]
        assert self.lib.h9r_set_soil(self.h, _i(st), *[_p(v) for v in a]) == 0

    @property
    def num_land(self):
        return int(self.lib.h9r_num_land(self.h))

    def land_index(self):
        out = np.zeros(self.num_land, np.int32)
        assert self.lib.h9r_get_land_index(self.h, _i(out)) == 0
        return out

    def init_state(self):
        assert self.lib.h9r_init_state(self.h) == 0

    def set_state(self, st, with_smp=True):
        nplants = np.ascontiguousarray(st.nplants, np.int32)
        a = [_f(getattr(st, n)) for n in ("h2osoi_liq", "zwt", "wa", "lai", "lai_litter",
                                           "plant_mass", "plant_foliage_mass", "plant_length",
                                           "rdepth", "rootr_col")]
        smp = _f(st.smp) if with_smp else None
        rc = self.lib.h9r_set_state(self.h, *[_p(v) for v in a], _i(nplants),
                                    _p(smp) if smp is not None else None)
        assert rc == 0

    def get_state(self, with_shared=False):
        from hybrid9_b200.state import H9State
        st = H9State.zeros(self.lat_c, self.lon_c)
        shared = np.zeros(8, np.float32)
        rc = self.lib.h9r_get_state(
            self.h, _p(st.h2osoi_liq), _p(st.zwt), _p(st.wa), _p(st.lai), _p(st.lai_litter),
            _p(st.plant_mass), _p(st.plant_foliage_mass), _p(st.plant_length), _p(st.rdepth),
            _p(st.rootr_col), _i(st.nplants), _p(st.smp), _p(shared))
        assert rc == 0
        return (st, shared) if with_shared else st

    def run_days(self, year_index, forcing):
        yi = np.ascontiguousarray(year_index, np.int32)
        a = [_f(forcing[k]) for k in FORCING]
        rc = self.lib.h9r_run_days(self.h, int(yi.shape[0]), _i(yi), *[_p(v) for v in a],
                                   self.per_cell_smp)
        assert rc >= 0, rc
        return rc

    def run_decade(self, idec_start, idec, forcing):
        a = [_f(forcing[k]) for k in FORCING]
        rc = self.lib.h9r_run_decade(self.h, idec_start, idec, int(a[0].shape[0]),
                                     *[_p(v) for v in a])
        assert rc >= 0, rc
        return rc

    def get_annual(self, iyr):
        s2 = (self.lat_c, self.lon_c)
        out = {k: np.zeros(s2, np.float32) for k in ("npp", "plant_mass", "rnf", "evap",
                                                      "theta_total")}
        out["theta"] = np.zeros(s2 + (8,), np.float32)
        rc = self.lib.h9r_get_annual(self.h, iyr, _p(out["npp"]), _p(out["plant_mass"]),
                                     _p(out["rnf"]), _p(out["evap"]), _p(out["theta_total"]),
                                     _p(out["theta"]))
        assert rc == 0
        return out

    def get_annual_forcing(self, iyr):
        out = np.zeros((7, self.lat_c, self.lon_c), np.float32)
        assert self.lib.h9r_get_annual_forcing(self.h, iyr, _p(out)) == 0
        return dict(zip(FORCING, out))

    def hydrology_step(self, forcing):
        s2 = (self.lat_c, self.lon_c)
        out = {"theta": np.zeros(s2 + (8,), np.float32)}
        for k in ("qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc", "w_imbalance"):
            out[k] = np.zeros(s2, np.float32)
        out["jwt"] = np.zeros(s2, np.int32)
        diag = np.zeros(s2 + (len(DIAG),), np.float32)
        a = [_f(forcing[k]) for k in FORCING]
        out["fault"] = self.lib.h9r_hydrology_step(
            self.h, *[_p(v) for v in a], _p(out["theta"]), _p(out["qflx_tran_veg_col"]),
            _p(out["qflx_evap_grnd"]), _p(out["rnf_inc"]), _p(out["w_imbalance"]),
            _i(out["jwt"]), _p(diag), self.per_cell_smp)
        assert out["fault"] >= 0
        out["diag"] = {k: diag[..., j] for j, k in enumerate(DIAG)}
        return out

    def grow_day(self, tas):
        s2 = (self.lat_c, self.lon_c)
        out = {k: np.zeros(s2, np.float32) for k in ("npp", "w_i", "fT")}
        assert self.lib.h9r_grow_day(self.h, _p(_f(tas)), _p(out["npp"]), _p(out["w_i"]),
                                     _p(out["fT"]), self.per_cell_smp) == 0
        return out

    def get_fault(self):
        code = C.c_uint32()
        x, y, day, sub, line = (C.c_int32() for _ in range(5))
        imb = C.c_float()
        assert self.lib.h9r_get_fault(self.h, C.byref(code), C.byref(x), C.byref(y),
                                      C.byref(day), C.byref(sub), C.byref(imb),
                                      C.byref(line)) == 0
        return dict(code=code.value, x=x.value, y=y.value, day=day.value, substep=sub.value,
                    imbalance=imb.value, line=line.value)

    def geometry(self):
        dz, zc = np.zeros(10, np.float32), np.zeros(10, np.float32)
        dt = C.c_float()
        assert self.lib.h9r_get_geometry(self.h, _p(dz), _p(zc), C.byref(dt)) == 0
        return dz, zc, dt.value

    def time_boy(self, year):
        return self.lib.h9r_time_boy(self.h, year)


def make_ref(world, nisurf=48, nyr=1, kind="ref", per_cell_smp=True):
    from hybrid9_b200 import synth
    r = RefModel(kind, per_cell_smp)
    r.configure(world.nx, world.ny, nisurf, synth.ZI_DRIVER, nyr=nyr)
    r.set_soil(world.soil_tex, world.theta_s, world.hksat, world.bsw, world.psi_s, world.fmax)
    return r


class RefRanks:
    """The reference's parallel mode on one host: `nranks` independent instances, each owning
    its own block of land cells and its own module state, stepped concurrently (one thread
    each; ctypes releases the GIL).  This is what MPI ranks are to PROGRAM H9 -- blocks never
    exchange anything during time stepping (INIT.f90:271-296, HYBRID9.f90:120-295).  Used by
    bench.py for the CPU baseline; `kind` "o3" is the -O3 build."""

    def __init__(self, world, forcing, ncell, nisurf, nranks, kind="o3"):
        from hybrid9_b200 import synth
        self.ranks, self.forcing, self.cells = [], [], 0
        nranks = max(1, min(nranks, ncell))
        for t in range(nranks):
            a, b = ncell * t // nranks, ncell * (t + 1) // nranks
            cw = synth.compact_world(world, b - a, start=a)
            r = make_ref(cw, nisurf=nisurf, nyr=1, kind=kind, per_cell_smp=False)
            r.init_state()
            self.ranks.append(r)
            self.forcing.append(synth.compact_forcing(world, forcing, b - a, start=a))
            self.cells += r.num_land

    def run_days(self, ndays):
        """every rank runs HYBRID9.f90:120-295 over the first `ndays` days; returns seconds"""
        import threading
        import time
        yi = np.ones(ndays, np.int32)
        fs = [{k: np.ascontiguousarray(v[:ndays]) for k, v in f.items()} for f in self.forcing]
        th = [threading.Thread(target=r.run_days, args=(yi, f)) for r, f in zip(self.ranks, fs)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    def close(self):
        for r in self.ranks:
            r.close()
