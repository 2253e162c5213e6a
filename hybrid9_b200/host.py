"""Python host layer over the C ABI of libh9gpu.so (include/h9gpu.h).

Mirrors, call for call, what the Fortran host does through ISO_C_BINDING
(INTEGRATION.md): arrays are handed over in the reference's own memory order
(numpy C-order (lat_c, lon_c, 8) == Fortran (8, lon_c, lat_c); forcing
(ndays, lat_c, lon_c) == Fortran (lon_c, lat_c, ndays)).  Nothing is computed
here: every method is one C-ABI call, and the library has no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .state import H9State

MATH_EXACT = 0
MATH_FAST = 1
OPT_REAL_EVAP = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FORCING = ("tas", "rlds", "rsds", "huss", "ps", "pr", "rhs")  # READ_PGF.f90 order

c_f = C.POINTER(C.c_float)
c_i = C.POINTER(C.c_int32)


class H9Error(RuntimeError):
    pass


class _Fault(C.Structure):
    _fields_ = [("any", C.c_uint32), ("code", C.c_uint32), ("x", C.c_int32), ("y", C.c_int32),
                ("day", C.c_int32), ("substep", C.c_int32), ("imbalance", C.c_float),
                ("n_faulted", C.c_int64)]


@dataclass
class H9Fault:
    any: int
    code: int
    x: int
    y: int
    day: int
    substep: int
    imbalance: float
    n_faulted: int


# name -> (restype, argtypes); every symbol include/h9gpu.h declares
ABI = {
    "h9_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "h9_destroy": (C.c_int, [C.c_void_p]),
    "h9_last_error": (C.c_char_p, [C.c_void_p]),
    "h9_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_f, C.c_int]),
    "h9_set_math": (C.c_int, [C.c_void_p, C.c_int]),
    "h9_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "h9_set_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "h9_set_soil": (C.c_int, [C.c_void_p, c_i, c_f, c_f, c_f, c_f, c_f]),
    "h9_num_land": (C.c_int64, [C.c_void_p]),
    "h9_get_land_index": (C.c_int, [C.c_void_p, c_i]),
    "h9_set_state": (C.c_int, [C.c_void_p] + [c_f] * 10 + [c_i, c_f]),
    "h9_get_state": (C.c_int, [C.c_void_p] + [c_f] * 10 + [c_i, c_f]),
    "h9_run_days": (C.c_int, [C.c_void_p, C.c_int, c_i] + [c_f] * 7),
    "h9_run_days_device": (C.c_int, [C.c_void_p, C.c_int, c_i, C.c_void_p, C.c_size_t, C.c_size_t]),
    "h9_pack_forcing": (C.c_int, [C.c_void_p, C.c_int] + [c_f] * 7 +
                        [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "h9_get_annual": (C.c_int, [C.c_void_p, C.c_int] + [c_f] * 6),
    "h9_annual_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]),
    "h9_get_fault": (C.c_int, [C.c_void_p, C.POINTER(_Fault)]),
    "h9_clear_fault": (C.c_int, [C.c_void_p]),
    "h9_stream": (C.c_void_p, [C.c_void_p]),
    "h9_synchronize": (C.c_int, [C.c_void_p]),
    "h9_host_alloc": (C.c_void_p, [C.c_size_t]),
    "h9_host_free": (None, [C.c_void_p]),
    "h9_launch_count": (C.c_int64, [C.c_void_p]),
    "h9_h2d_bytes": (C.c_int64, [C.c_void_p]),
    "h9_d2h_bytes": (C.c_int64, [C.c_void_p]),
    "h9_step_kernel_ms": (C.c_double, [C.c_void_p]),
    "h9_reset_counters": (C.c_int, [C.c_void_p]),
    "h9_hydrology_step": (C.c_int, [C.c_void_p] + [c_f] * 7 + [c_f] * 5 + [c_i]),
    "h9_grow_day": (C.c_int, [C.c_void_p, c_f, c_f, c_f, c_f]),
    "h9_regrid_soil_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [c_f] * 8),
    "h9_partition_lat_bands": (C.c_int, [C.c_int, C.c_int, c_i, c_f, C.c_int, c_i, c_i,
                                         C.POINTER(C.c_int64)]),
    "h9_kernel_variant": (C.c_char_p, [C.c_void_p]),
    "h9_comm_unique_id": (C.c_int, [C.c_void_p]),
    "h9_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "h9_comm_destroy": (C.c_int, [C.c_void_p]),
    "h9_comm_land_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "h9_annual_collective": (C.c_int, [C.c_void_p, C.c_int]),
    "h9_get_gathered_annual": (C.c_int, [C.c_void_p, C.c_int, c_f]),
    "h9_get_budget": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "h9_gathered_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
}

COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """h9_comm_unique_id: rank 0 creates the NCCL id; the host broadcasts these bytes
    (MPI_Bcast in the Fortran host, torch.distributed / a file in the tests)."""
    lib = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib.h9_comm_unique_id(buf)
    if rc != 0:
        raise H9Error(f"h9_comm_unique_id failed ({rc}): libnccl not loadable (set H9_NCCL_LIB)")
    return buf.raw


def library_path() -> str:
    return os.environ.get("H9GPU_LIB", os.path.join(_HERE, "libh9gpu.so"))


def load_library() -> C.CDLL:
    """dlopen libh9gpu.so and bind every symbol of include/h9gpu.h.  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise H9Error(
            f"{path} is missing: build it with `python -m hybrid9_b200.build` "
            "(there is no CPU fallback for the HYDROLOGY/GROW path)")
    lib = C.CDLL(path)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _f(a):
    return None if a is None else a.ctypes.data_as(c_f)


def _i(a):
    return None if a is None else a.ctypes.data_as(c_i)


def _chk32(a, shape, name, dtype=np.float32):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise H9Error(f"{name}: need a C-contiguous {np.dtype(dtype).name} array")
    if tuple(a.shape) != tuple(shape):
        raise H9Error(f"{name}: shape {a.shape} != expected {tuple(shape)}")
    return a


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array over page-locked host memory from h9_host_alloc (never freed: bench lifetime)."""
    lib = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.h9_host_alloc(max(n, 1))
    if not p:
        raise H9Error("h9_host_alloc failed")
    buf = (C.c_char * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def partition_lat_bands(soil_tex: np.ndarray, theta_s: np.ndarray, nranks: int):
    """Balanced contiguous latitude bands: (lat_s[1-based], lat_count, n_land) per rank."""
    lib = load_library()
    lat_c, lon_c = soil_tex.shape
    ls = np.zeros(nranks, np.int32)
    lc = np.zeros(nranks, np.int32)
    nl = np.zeros(nranks, np.int64)
    rc = lib.h9_partition_lat_bands(lon_c, lat_c, _i(np.ascontiguousarray(soil_tex, np.int32)),
                                    _f(np.ascontiguousarray(theta_s, np.float32)), nranks,
                                    _i(ls), _i(lc), nl.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        raise H9Error(f"h9_partition_lat_bands failed ({rc})")
    return ls, lc, nl


class H9:
    """One GPU context == one block of the reference's domain decomposition."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.h9_create(C.byref(h), device)
        if rc != 0 or not h:
            raise H9Error(f"h9_create failed ({rc}): no usable CUDA device; there is no CPU path")
        self.h = h
        self.lon_c = self.lat_c = self.nisurf = self.nyr = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.h9_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str) -> int:
        if rc < 0:
            msg = self.lib.h9_last_error(self.h)
            raise H9Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
        return rc

    # -- configuration -------------------------------------------------------
    def configure(self, lon_c: int, lat_c: int, nisurf: int, zi, nyr: int = 1):
        zi = np.ascontiguousarray(zi, dtype=np.float32)
        if zi.shape != (10,):
            raise H9Error("zi must hold zi(0:9)")
        self._ck(self.lib.h9_configure(self.h, lon_c, lat_c, nisurf, _f(zi), nyr), "h9_configure")
        self.lon_c, self.lat_c, self.nisurf, self.nyr = lon_c, lat_c, nisurf, nyr

    def set_math(self, mode: int):
        self._ck(self.lib.h9_set_math(self.h, mode), "h9_set_math")

    def set_option(self, option: int, value: int):
        self._ck(self.lib.h9_set_option(self.h, option, value), "h9_set_option")

    def set_tuning(self, tile_days: int = 0, block: int = 0):
        self._ck(self.lib.h9_set_tuning(self.h, tile_days, block), "h9_set_tuning")

    def kernel_variant(self) -> str:
        v = self.lib.h9_kernel_variant(self.h)
        return v.decode() if v else ""

    def set_soil(self, soil_tex, theta_s, hksat, bsw, psi_s, fmax):
        s2, s3 = (self.lat_c, self.lon_c), (self.lat_c, self.lon_c, 8)
        _chk32(soil_tex, s2, "soil_tex", np.int32)
        for n, a in (("theta_s", theta_s), ("hksat", hksat), ("bsw", bsw), ("psi_s", psi_s)):
            _chk32(a, s3, n)
        _chk32(fmax, s2, "fmax")
        self._ck(self.lib.h9_set_soil(self.h, _i(soil_tex), _f(theta_s), _f(hksat), _f(bsw),
                                      _f(psi_s), _f(fmax)), "h9_set_soil")

    @property
    def num_land(self) -> int:
        return int(self.lib.h9_num_land(self.h))

    def land_index(self) -> np.ndarray:
        out = np.zeros(max(self.num_land, 0), np.int32)
        self._ck(self.lib.h9_get_land_index(self.h, _i(out)), "h9_get_land_index")
        return out

    # -- state -----------------------------------------------------------------
    def set_state(self, st: H9State, with_smp: bool = True):
        s2 = (self.lat_c, self.lon_c)
        _chk32(st.h2osoi_liq, s2 + (8,), "h2osoi_liq")
        _chk32(st.rootr_col, s2 + (9,), "rootr_col")
        _chk32(st.nplants, s2, "nplants", np.int32)
        self._ck(self.lib.h9_set_state(
            self.h, _f(st.h2osoi_liq), _f(st.zwt), _f(st.wa), _f(st.lai), _f(st.lai_litter),
            _f(st.plant_mass), _f(st.plant_foliage_mass), _f(st.plant_length), _f(st.rdepth),
            _f(st.rootr_col), _i(st.nplants), _f(st.smp) if with_smp else None), "h9_set_state")

    def get_state(self) -> H9State:
        st = H9State.zeros(self.lat_c, self.lon_c)
        self._ck(self.lib.h9_get_state(
            self.h, _f(st.h2osoi_liq), _f(st.zwt), _f(st.wa), _f(st.lai), _f(st.lai_litter),
            _f(st.plant_mass), _f(st.plant_foliage_mass), _f(st.plant_length), _f(st.rdepth),
            _f(st.rootr_col), _i(st.nplants), _f(st.smp)), "h9_get_state")
        return st

    # -- hot path ---------------------------------------------------------------
    def _forcing_args(self, forcing: dict, shape):
        return [_f(_chk32(forcing[k], shape, k)) for k in FORCING]

    def run_days(self, year_index, forcing: dict) -> int:
        """h9_run_days: forcing[k] of shape (ndays, lat_c, lon_c); returns the fault bits (0 = ok)."""
        yi = np.ascontiguousarray(year_index, np.int32)
        nd = int(yi.shape[0])
        args = self._forcing_args(forcing, (nd, self.lat_c, self.lon_c))
        return self._ck(self.lib.h9_run_days(self.h, nd, _i(yi), *args), "h9_run_days")

    def pack_forcing(self, forcing: dict, ndays: int):
        args = self._forcing_args(forcing, (ndays, self.lat_c, self.lon_c))
        p, ds, ps = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._ck(self.lib.h9_pack_forcing(self.h, ndays, *args, C.byref(p), C.byref(ds),
                                          C.byref(ps)), "h9_pack_forcing")
        return p.value, ds.value, ps.value

    def run_days_device(self, year_index, d_forcing: int, day_stride: int, plane_stride: int) -> int:
        yi = np.ascontiguousarray(year_index, np.int32)
        return self._ck(self.lib.h9_run_days_device(self.h, int(yi.shape[0]), _i(yi),
                                                    C.c_void_p(d_forcing), day_stride,
                                                    plane_stride), "h9_run_days_device")

    def get_annual(self, iyr: int, fill=np.nan) -> dict:
        """Annual means of year iyr on the (lat_c, lon_c) grid; non-land keeps the
        reference's fill (NaN; 0 for theta_total, INIT.f90:402-414)."""
        s2 = (self.lat_c, self.lon_c)
        out = {k: np.full(s2, fill, np.float32) for k in ("npp", "plant_mass", "rnf", "evap")}
        out["theta_total"] = np.zeros(s2, np.float32)
        out["theta"] = np.full(s2 + (8,), fill, np.float32)
        self._ck(self.lib.h9_get_annual(self.h, iyr, _f(out["npp"]), _f(out["plant_mass"]),
                                        _f(out["rnf"]), _f(out["evap"]), _f(out["theta_total"]),
                                        _f(out["theta"])), "h9_get_annual")
        return out

    def annual_device(self, iyr: int, budget: bool = True):
        p, ps, b = C.c_void_p(), C.c_size_t(), C.c_void_p()
        self._ck(self.lib.h9_annual_device(self.h, iyr, C.byref(p), C.byref(ps),
                                           C.byref(b) if budget else None), "h9_annual_device")
        return p.value, ps.value, b.value

    def hydrology_step(self, forcing: dict) -> dict:
        """One HYDROLOGY call for all land cells; forcing[k] of shape (lat_c, lon_c)."""
        s2 = (self.lat_c, self.lon_c)
        args = self._forcing_args(forcing, s2)
        out = {"theta": np.zeros(s2 + (8,), np.float32)}
        for k in ("qflx_tran_veg_col", "qflx_evap_grnd", "rnf_inc", "w_imbalance"):
            out[k] = np.zeros(s2, np.float32)
        out["jwt"] = np.zeros(s2, np.int32)
        out["fault"] = self._ck(self.lib.h9_hydrology_step(
            self.h, *args, _f(out["theta"]), _f(out["qflx_tran_veg_col"]),
            _f(out["qflx_evap_grnd"]), _f(out["rnf_inc"]), _f(out["w_imbalance"]),
            _i(out["jwt"])), "h9_hydrology_step")
        return out

    def grow_day(self, tas) -> dict:
        s2 = (self.lat_c, self.lon_c)
        _chk32(tas, s2, "tas")
        out = {k: np.zeros(s2, np.float32) for k in ("npp", "w_i", "fT")}
        self._ck(self.lib.h9_grow_day(self.h, _f(tas), _f(out["npp"]), _f(out["w_i"]),
                                      _f(out["fT"])), "h9_grow_day")
        return out

    def regrid_soil_layer(self, lon_c, lat_c, layer, theta_s_in, k_s_in, lambda_in, psi_s_in,
                          theta_s, hksat, bsw, psi_s):
        """INIT.f90:573-633 for one layer: 30-arc-second fields (lat_c*60, lon_c*60) -> element
        `layer` of the (lat_c, lon_c, 8) arrays."""
        fine, coarse = (lat_c * 60, lon_c * 60), (lat_c, lon_c, 8)
        ins = [_f(_chk32(a, fine, "fine field")) for a in (theta_s_in, k_s_in, lambda_in, psi_s_in)]
        outs = [_f(_chk32(a, coarse, "soil array")) for a in (theta_s, hksat, bsw, psi_s)]
        self._ck(self.lib.h9_regrid_soil_layer(self.h, lon_c, lat_c, layer, *ins, *outs),
                 "h9_regrid_soil_layer")

    # -- multi-GPU: NCCL behind the C ABI ----------------------------------------
    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        if len(unique_id) != COMM_ID_BYTES:
            raise H9Error("unique_id must be COMM_ID_BYTES long")
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES)
        self._ck(self.lib.h9_comm_init(self.h, nranks, rank, buf), "h9_comm_init")
        self.nranks, self.rank = nranks, rank

    def comm_destroy(self):
        self._ck(self.lib.h9_comm_destroy(self.h), "h9_comm_destroy")

    def comm_land_counts(self) -> np.ndarray:
        out = np.zeros(self.nranks, np.int64)
        self._ck(self.lib.h9_comm_land_counts(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))),
                 "h9_comm_land_counts")
        return out

    def annual_collective(self, iyr: int):
        """Budget kernel + FP64 all-reduce + ragged all-gather of the annual planes, enqueued
        on the ctx's stream behind the stepping kernel (no host synchronisation)."""
        self._ck(self.lib.h9_annual_collective(self.h, iyr), "h9_annual_collective")

    def get_gathered_annual(self, r: int, n_land_r: int) -> np.ndarray:
        out = np.zeros((13, int(n_land_r)), np.float32)
        self._ck(self.lib.h9_get_gathered_annual(self.h, r, _f(out)), "h9_get_gathered_annual")
        return out

    def collective_fence(self):
        """Order later work on the ctx's stream behind the last h9_annual_collective (which may
        run on the library's communication stream); no host synchronisation."""
        self._ck(self.lib.h9_gathered_device(self.h, None, None), "h9_gathered_device")

    def get_budget(self, iyr: int) -> np.ndarray:
        out = np.zeros(8, np.float64)
        self._ck(self.lib.h9_get_budget(self.h, iyr, out.ctypes.data_as(C.POINTER(C.c_double))),
                 "h9_get_budget")
        return out

    # -- faults, sync, counters ---------------------------------------------------
    def get_fault(self) -> H9Fault:
        f = _Fault()
        self._ck(self.lib.h9_get_fault(self.h, C.byref(f)), "h9_get_fault")
        return H9Fault(f.any, f.code, f.x, f.y, f.day, f.substep, f.imbalance, f.n_faulted)

    def clear_fault(self):
        self._ck(self.lib.h9_clear_fault(self.h), "h9_clear_fault")

    def synchronize(self):
        self._ck(self.lib.h9_synchronize(self.h), "h9_synchronize")

    @property
    def stream(self) -> int:
        return int(self.lib.h9_stream(self.h) or 0)

    def counters(self) -> dict:
        return {"launches": int(self.lib.h9_launch_count(self.h)),
                "h2d_bytes": int(self.lib.h9_h2d_bytes(self.h)),
                "d2h_bytes": int(self.lib.h9_d2h_bytes(self.h)),
                "step_kernel_ms": float(self.lib.h9_step_kernel_ms(self.h))}

    def reset_counters(self):
        self._ck(self.lib.h9_reset_counters(self.h), "h9_reset_counters")
