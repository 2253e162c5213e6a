"""ctypes wrapper of the CPU oracle (oracle/h9_oracle.h) and of the host twin of
the kernel source (tests/twin).  TEST INFRASTRUCTURE: imported only by tests/,
bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke()."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
TWIN_DIR = os.path.join(ROOT, "tests", "twin")
_LIBS = {}
FORCING = ("tas", "rlds", "rsds", "huss", "ps", "pr", "rhs")


class StepDiag(C.Structure):
    pass


def _diag_fields(real):
    return [("theta", real * 8), ("qflx_tran_veg_col", real), ("qflx_evap_grnd", real),
            ("qflx_surf", real), ("rsub_top", real), ("qflx_rsub_sat", real), ("qflx_infl", real),
            ("qcharge", real), ("fsat", real), ("beta", real), ("rsc", real), ("w0", real),
            ("w1", real), ("rnf_inc", real), ("jwt_soilwater", C.c_int32),
            ("jwt_final", C.c_int32), ("fault", C.c_uint32)]


def build(which: str = "all"):
    """make the oracle libraries (and the twin) if sources are newer; needs only g++."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    subprocess.run(["make", "-s", "-C", TWIN_DIR], check=True)


def load(kind: str = "f32"):
    """kind: f32 (strict checker), f64 (noise floor), o3 (CPU baseline build)."""
    if kind in _LIBS:
        return _LIBS[kind]
    name = {"f32": "libh9oracle.so", "f64": "libh9oracle_f64.so", "o3": "libh9oracle_o3.so"}[kind]
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    real = C.c_double if kind == "f64" else C.c_float
    assert lib.h9o_sizeof_real() == C.sizeof(real)
    rp = C.POINTER(real)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    sig = {
        "h9o_create": (C.c_int, [C.POINTER(vp)]),
        "h9o_destroy": (C.c_int, [vp]),
        "h9o_configure": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, rp, C.c_int]),
        "h9o_set_soil": (C.c_int, [vp, ip, rp, rp, rp, rp, rp]),
        "h9o_num_land": (C.c_int64, [vp]),
        "h9o_get_land_index": (C.c_int, [vp, ip]),
        "h9o_init_state": (C.c_int, [vp]),
        "h9o_set_state": (C.c_int, [vp] + [rp] * 10 + [ip, rp]),
        "h9o_get_state": (C.c_int, [vp] + [rp] * 10 + [ip, rp]),
        "h9o_set_options": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "h9o_set_real_evap": (C.c_int, [vp, C.c_int]),
        "h9o_run_days": (C.c_int, [vp, C.c_int, ip] + [rp] * 7),
        "h9o_get_annual": (C.c_int, [vp, C.c_int] + [rp] * 6),
        "h9o_get_fault": (C.c_int, [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), ip, ip, ip,
                                    ip, rp, C.POINTER(C.c_int64)]),
        "h9o_clear_fault": (C.c_int, [vp]),
        "h9o_hydrology_step": (C.c_int, [vp] + [rp] * 7 + [rp] * 5 + [ip]),
        "h9o_grow_day": (C.c_int, [vp, rp, rp, rp, rp]),
        "h9o_last_step_diag": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "h9o_get_geometry": (C.c_int, [vp, rp, rp, rp]),
        "h9o_time_boy": (C.c_int, [C.c_int]),
        "h9o_regrid_soil_layer": (C.c_int, [C.c_int, C.c_int, C.c_int] + [rp] * 8),
    }
    for n, (res, args) in sig.items():
        fn = getattr(lib, n)
        fn.restype, fn.argtypes = res, args
    lib._real = real
    lib._np = np.float64 if kind == "f64" else np.float32
    _LIBS[kind] = lib
    return lib


class Oracle:
    """Same method names and array conventions as hybrid9_b200.host.H9."""

    def __init__(self, kind: str = "f32"):
        self.lib = load(kind)
        self.dt = self.lib._np
        self.rp = C.POINTER(self.lib._real)
        h = C.c_void_p()
        assert self.lib.h9o_create(C.byref(h)) == 0
        self.h = h
        self._keep = []

    def close(self):
        if self.h:
            self.lib.h9o_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _r(self, a):
        if a is None:
            return None
        b = np.ascontiguousarray(a, dtype=self.dt)
        self._keep.append(b)
        return b.ctypes.data_as(self.rp)

    @staticmethod
    def _i(a):
        return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))

    def configure(self, lon_c, lat_c, nisurf, zi, nyr=1):
        self.lon_c, self.lat_c, self.nisurf, self.nyr = lon_c, lat_c, nisurf, nyr
        zi = np.ascontiguousarray(zi, dtype=self.dt)
        rc = self.lib.h9o_configure(self.h, lon_c, lat_c, nisurf, zi.ctypes.data_as(self.rp), nyr)
        assert rc == 0

    def set_soil(self, soil_tex, theta_s, hksat, bsw, psi_s, fmax):
        st = np.ascontiguousarray(soil_tex, np.int32)
        rc = self.lib.h9o_set_soil(self.h, self._i(st), self._r(theta_s), self._r(hksat),
                                   self._r(bsw), self._r(psi_s), self._r(fmax))
        assert rc == 0
        self._keep.clear()

    @property
    def num_land(self):
        return int(self.lib.h9o_num_land(self.h))

    def land_index(self):
        out = np.zeros(self.num_land, np.int32)
        assert self.lib.h9o_get_land_index(self.h, self._i(out)) == 0
        return out

    def set_options(self, loop_order=0, smp_leak=0, nthreads=1):
        assert self.lib.h9o_set_options(self.h, loop_order, smp_leak, nthreads) == 0

    def set_real_evap(self, on):
        assert self.lib.h9o_set_real_evap(self.h, 1 if on else 0) == 0

    def init_state(self):
        assert self.lib.h9o_init_state(self.h) == 0

    def set_state(self, st, with_smp=True):
        nplants = np.ascontiguousarray(st.nplants, np.int32)
        rc = self.lib.h9o_set_state(
            self.h, self._r(st.h2osoi_liq), self._r(st.zwt), self._r(st.wa), self._r(st.lai),
            self._r(st.lai_litter), self._r(st.plant_mass), self._r(st.plant_foliage_mass),
            self._r(st.plant_length), self._r(st.rdepth), self._r(st.rootr_col),
            self._i(nplants), self._r(st.smp) if with_smp else None)
        assert rc == 0
        self._keep.clear()

    def get_state(self):
        from hybrid9_b200.state import H9State
        st = H9State.zeros(self.lat_c, self.lon_c)
        if self.dt is not np.float32:
            for n in st.names():
                if n != "nplants":
                    setattr(st, n, getattr(st, n).astype(self.dt))
        p = lambda a: a.ctypes.data_as(self.rp)  # noqa: E731
        rc = self.lib.h9o_get_state(
            self.h, p(st.h2osoi_liq), p(st.zwt), p(st.wa), p(st.lai), p(st.lai_litter),
            p(st.plant_mass), p(st.plant_foliage_mass), p(st.plant_length), p(st.rdepth),
            p(st.rootr_col), self._i(st.nplants), p(st.smp))
        assert rc == 0
        return st

    def run_days(self, year_index, forcing):
        yi = np.ascontiguousarray(year_index, np.int32)
        args = [self._r(forcing[k]) for k in FORCING]
        rc = self.lib.h9o_run_days(self.h, int(yi.shape[0]), self._i(yi), *args)
        self._keep.clear()
        assert rc >= 0
        return rc

    def get_annual(self, iyr, fill=np.nan):
        s2 = (self.lat_c, self.lon_c)
        out = {k: np.full(s2, fill, self.dt) for k in ("npp", "plant_mass", "rnf", "evap")}
        out["theta_total"] = np.zeros(s2, self.dt)
        out["theta"] = np.full(s2 + (8,), fill, self.dt)
        p = lambda a: a.ctypes.data_as(self.rp)  # noqa: E731
        rc = self.lib.h9o_get_annual(self.h, iyr, p(out["npp"]), p(out["plant_mass"]),
                                     p(out["rnf"]), p(out["evap"]), p(out["theta_total"]),
                                     p(out["theta"]))
        assert rc == 0
        return out

    def hydrology_step(self, forcing):
        s2 = (self.lat_c, self.lon_c)
        out = {"theta": np.zeros(s2 + (8,), self.dt)}
        for k in ("qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc", "w_imbalance"):
            out[k] = np.zeros(s2, self.dt)
        out["jwt"] = np.zeros(s2, np.int32)
        p = lambda a: a.ctypes.data_as(self.rp)  # noqa: E731
        args = [self._r(forcing[k]) for k in FORCING]
        out["fault"] = self.lib.h9o_hydrology_step(
            self.h, *args, p(out["theta"]), p(out["qflx_tran_veg_col"]), p(out["qflx_evap_grnd"]),
            p(out["rnf_inc"]), p(out["w_imbalance"]), self._i(out["jwt"]))
        self._keep.clear()
        return out

    def grow_day(self, tas):
        s2 = (self.lat_c, self.lon_c)
        out = {k: np.zeros(s2, self.dt) for k in ("npp", "w_i", "fT")}
        p = lambda a: a.ctypes.data_as(self.rp)  # noqa: E731
        assert self.lib.h9o_grow_day(self.h, self._r(tas), p(out["npp"]), p(out["w_i"]),
                                     p(out["fT"])) == 0
        self._keep.clear()
        return out

    def step_diag(self, x, y):
        class D(C.Structure):
            _fields_ = _diag_fields(self.lib._real)
        d = D()
        assert self.lib.h9o_last_step_diag(self.h, x, y, C.byref(d)) == 0
        return d

    def get_fault(self):
        any_, code = C.c_uint32(), C.c_uint32()
        x, y, day, sub = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        imb, nf = self.lib._real(), C.c_int64()
        rc = self.lib.h9o_get_fault(self.h, C.byref(any_), C.byref(code), C.byref(x), C.byref(y),
                                    C.byref(day), C.byref(sub), C.byref(imb), C.byref(nf))
        assert rc == 0
        return dict(any=any_.value, code=code.value, x=x.value, y=y.value, day=day.value,
                    substep=sub.value, imbalance=imb.value, n_faulted=nf.value)

    def geometry(self):
        dz = np.zeros(10, self.dt)
        zc = np.zeros(10, self.dt)
        dt = self.lib._real()
        p = lambda a: a.ctypes.data_as(self.rp)  # noqa: E731
        assert self.lib.h9o_get_geometry(self.h, p(dz), p(zc), C.byref(dt)) == 0
        return dz, zc, dt.value


def load_twin():
    path = os.path.join(TWIN_DIR, "libh9twin.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    lib.h9t_run.restype = C.c_int
    lib.h9t_run.argtypes = ([C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp] + [fp] * 16 + [ip, fp, fp, C.c_int, up,
                            fp, fp, fp, fp, ip, fp, fp, fp])
    for n, a in (("h9t_pow", [C.c_float, C.c_float]), ("h9t_exp", [C.c_float]), ("h9t_log", [C.c_float])):
        getattr(lib, n).restype = C.c_float
        getattr(lib, n).argtypes = a
    return lib


def twin_run(world, state, forcing, nisurf, zi, do_grow=True, math="libm", nsteps=-1):
    """Run the host twin of the kernel source over the land cells of `world`.
    forcing[k]: (ndays, ny, nx).  nsteps: HYDROLOGY calls per day (-1 = nisurf; dt is
    always 86400/nisurf).  Returns (H9State, extras dict)."""
    lib = load_twin()
    land = world.land
    yy, xx = np.nonzero(land)
    n = yy.size
    nd = forcing["tas"].shape[0]
    f = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
    comp = {k: f(getattr(state, k)[yy, xx]) for k in
            ("h2osoi_liq", "smp", "zwt", "wa", "lai", "lai_litter")}
    rootr = f(state.rootr_col[yy, xx, :8])
    pk = {k: f(getattr(state, k)[yy, xx, 0]) for k in
          ("plant_mass", "plant_foliage_mass", "plant_length", "rdepth")}
    nplants = np.ascontiguousarray(state.nplants[yy, xx], np.int32)
    par = {k: f(getattr(world, k)[yy, xx]) for k in ("theta_s", "hksat", "bsw", "psi_s", "fmax")}
    forc = np.zeros((nd, 7, n), np.float32)
    for j, k in enumerate(FORCING):
        forc[:, j, :] = forcing[k][:, yy, xx]
    rnf = np.zeros(n, np.float32)
    fault = np.zeros(n, np.uint32)
    theta = np.zeros((n, 8), np.float32)
    ltran, levap, limb = (np.zeros(n, np.float32) for _ in range(3))
    ljwt = np.zeros(n, np.int32)
    dn, dw, dft = (np.zeros((nd, n), np.float32) for _ in range(3))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    zi = f(zi)
    rc = lib.h9t_run(0 if math == "libm" else 1, nsteps, n, nd, nisurf, P(zi), P(comp["h2osoi_liq"]), P(comp["smp"]), P(rootr),
                     P(par["theta_s"]), P(par["hksat"]), P(par["bsw"]), P(par["psi_s"]),
                     P(par["fmax"]), P(comp["zwt"]), P(comp["wa"]), P(comp["lai"]),
                     P(comp["lai_litter"]), P(pk["plant_mass"]), P(pk["plant_foliage_mass"]),
                     P(pk["plant_length"]), P(pk["rdepth"]),
                     nplants.ctypes.data_as(C.POINTER(C.c_int32)), P(rnf), P(forc),
                     1 if do_grow else 0, fault.ctypes.data_as(C.POINTER(C.c_uint32)), P(theta),
                     P(ltran), P(levap), P(limb), ljwt.ctypes.data_as(C.POINTER(C.c_int32)),
                     P(dn), P(dw), P(dft))
    assert rc == 0
    out = state.copy()
    out.h2osoi_liq[yy, xx] = comp["h2osoi_liq"]
    out.smp[yy, xx] = comp["smp"]
    out.rootr_col[yy, xx, :8] = rootr
    for k in ("zwt", "wa", "lai", "lai_litter"):
        getattr(out, k)[yy, xx] = comp[k]
    for k in pk:
        getattr(out, k)[yy, xx, 0] = pk[k]
    extras = dict(rnf_sum=rnf, fault=fault, theta=theta, qflx_tran_veg_col=ltran,
                  qflx_evap_grnd=levap, w_imbalance=limb, jwt=ljwt, npp=dn, w_i=dw, fT=dft,
                  yy=yy, xx=xx)
    return out, extras
