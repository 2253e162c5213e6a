"""Host-side state container and the reference's initial state.

``H9State`` holds the arrays the Fortran host owns in module SHARED
(SHARED.f90:30-101,198,459-472) in the reference's memory order: a numpy C-order
array of shape (lat_c, lon_c, 8) is byte-identical to Fortran (8, lon_c, lat_c).
``init_state`` is what INIT.f90:707-811 computes for every land cell, in float32
and in the same operation order; it is host code (INIT stays on the host), not
part of the GPU path.
"""
from __future__ import annotations

from dataclasses import dataclass, fields

import numpy as np

F32 = np.float32
FIELDS_8 = ("h2osoi_liq", "smp")
FIELDS_1 = ("zwt", "wa", "lai", "lai_litter")
FIELDS_K = ("plant_mass", "plant_foliage_mass", "plant_length", "rdepth")  # (nplants_max=1,x,y)


@dataclass
class H9State:
    h2osoi_liq: np.ndarray          # (lat_c, lon_c, 8)  mm
    zwt: np.ndarray                 # (lat_c, lon_c)     m
    wa: np.ndarray                  # (lat_c, lon_c)     mm
    lai: np.ndarray                 # (lat_c, lon_c)
    lai_litter: np.ndarray          # (lat_c, lon_c)
    plant_mass: np.ndarray          # (lat_c, lon_c, 1)
    plant_foliage_mass: np.ndarray  # (lat_c, lon_c, 1)
    plant_length: np.ndarray        # (lat_c, lon_c, 1)
    rdepth: np.ndarray              # (lat_c, lon_c, 1)
    rootr_col: np.ndarray           # (lat_c, lon_c, 9)
    nplants: np.ndarray             # (lat_c, lon_c) int32
    smp: np.ndarray                 # (lat_c, lon_c, 8)  mm (module scratch in the reference)

    @staticmethod
    def zeros(lat_c: int, lon_c: int) -> "H9State":
        z = lambda *s: np.zeros((lat_c, lon_c) + s, dtype=F32)  # noqa: E731
        return H9State(z(8), z(), z(), z(), z(), z(1), z(1), z(1), z(1), z(9),
                       np.zeros((lat_c, lon_c), dtype=np.int32), z(8))

    def copy(self) -> "H9State":
        return H9State(**{f.name: getattr(self, f.name).copy() for f in fields(self)})

    def names(self):
        return [f.name for f in fields(self)]


def land_mask(soil_tex: np.ndarray, theta_s: np.ndarray) -> np.ndarray:
    """Land predicate of HYBRID9.f90:122-123 / INIT.f90:719-720 (float32 sum in layer order)."""
    s = np.zeros(soil_tex.shape, dtype=F32)
    for i in range(8):
        s = (s + theta_s[..., i].astype(F32)).astype(F32)
    return (soil_tex > 0) & (soil_tex != 13) & (s > F32(1.0e-8))


def geometry(zi: np.ndarray, nisurf: int):
    """dt, dz(1:9), zc(1:9) of INIT.f90:214,252-257 (index 0 unused), float32."""
    zi = np.asarray(zi, dtype=F32)
    dz = np.zeros(10, dtype=F32)
    zc = np.zeros(10, dtype=F32)
    dz[1:] = zi[1:] - zi[:-1]
    zc[1:] = zi[1:] - dz[1:] / F32(2.0)
    dt = F32(86400.0) / F32(nisurf)
    return dt, dz, zc


def init_state(soil_tex: np.ndarray, theta_s: np.ndarray, zi: np.ndarray) -> H9State:
    """INIT.f90:707-811 for every land cell; zeros elsewhere."""
    lat_c, lon_c = soil_tex.shape
    st = H9State.zeros(lat_c, lon_c)
    land = land_mask(soil_tex, theta_s)
    zi = np.asarray(zi, dtype=F32)
    _, dz, _ = geometry(zi, 48)
    rhow = F32(1000.0)
    for i in range(8):  # :730-731
        v = F32(0.4) * theta_s[..., i].astype(F32)
        v = (v * dz[i + 1]).astype(F32)
        v = (v * rhow).astype(F32)
        v = (v / F32(1000.0)).astype(F32)
        st.h2osoi_liq[..., i] = np.where(land, v, F32(0))
    st.zwt[land] = (zi[8] + F32(5000.0)) / F32(1000.0)  # :739
    st.wa[land] = F32(4000.0)                            # :744
    st.lai_litter[land] = F32(0.001)                     # :748
    st.nplants[land] = 1                                 # :752
    plant_mass = F32(1.0)                                # :770
    pfm = F32(0.0435)                                    # :771
    base = F32(F32(400.0) * plant_mass) / F32(3.142e-3)
    plant_length = F32(np.power(F32(base), F32(F32(1.0) / F32(3.0))))  # :776
    lai = F32(F32(0.0) + F32(pfm * F32(23.0e-3)) / F32(1.0))           # :781 (sla INIT.f90:154)
    rdepth = F32(F32(0.3) * plant_length)                              # :786
    decay = F32(np.exp(F32(np.log(F32(0.1))) / F32(rdepth / F32(10.0))))  # :791
    rootr = np.zeros(9, dtype=F32)
    for i in range(1, 9):  # :793-797
        a = F32(F32(1.0) - F32(np.power(decay, F32(zi[i] / F32(10.0)))))
        b = F32(F32(1.0) - F32(np.power(decay, F32(zi[i - 1] / F32(10.0)))))
        rootr[i - 1] = F32(F32(rootr[i - 1] + a) - b)
    st.plant_mass[land, 0] = plant_mass
    st.plant_foliage_mass[land, 0] = pfm
    st.plant_length[land, 0] = plant_length
    st.lai[land] = lai
    st.rdepth[land, 0] = rdepth
    st.rootr_col[land, :] = rootr
    return st
