"""INIT's 30-arc-second -> half-degree soil pre-processing (INIT.f90:573-633).
The GPU kernel adds the 60x60 fine cells in the reference's own order, so its float sums must
equal the oracle's bit for bit; the oracle is checked against a float64 block mean."""
import ctypes as C

import numpy as np
import pytest

import oracle_py
from hybrid9_b200 import H9

LON_C, LAT_C = 7, 5


def fine_fields(seed=3):
    rng = np.random.default_rng(seed)
    shp = (LAT_C * 60, LON_C * 60)
    ts = rng.uniform(300, 550, shp).astype(np.float32)      # 0.001 cm^3/cm^3
    ks = np.exp(rng.uniform(np.log(0.5), np.log(488), shp)).astype(np.float32)   # cm/day
    lm = rng.uniform(80, 350, shp).astype(np.float32)       # 0.001 unitless
    ps = -rng.uniform(5, 80, shp).astype(np.float32)        # cm
    ts[rng.random(shp) < 0.2] = -9999.0                     # missing / ocean fine cells
    ts[0:60, 60:120] = -9999.0                              # a half-degree cell with no data at all
    return ts, ks, lm, ps


def oracle_regrid(fields, layer):
    lib = oracle_py.load("f32")
    out = [np.zeros((LAT_C, LON_C, 8), np.float32) for _ in range(4)]
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    assert lib.h9o_regrid_soil_layer(LON_C, LAT_C, layer, *[p(a) for a in fields], *[p(a) for a in out]) == 0
    return out


def test_oracle_regrid_matches_a_float64_block_mean():
    fields = fine_fields()
    ths, hks, bsw, psi = oracle_regrid(fields, 3)
    ts = fields[0].astype(np.float64).reshape(LAT_C, 60, LON_C, 60)
    ok = ts >= 0
    n = ok.sum(axis=(1, 3))
    for f, out, conv in ((0, ths, lambda m: m / 1e3), (1, hks, lambda m: 10 * m / 86400),
                         (3, psi, lambda m: 10 * m)):
        v = fields[f].astype(np.float64).reshape(LAT_C, 60, LON_C, 60)
        mean = np.where(n > 0, (v * ok).sum(axis=(1, 3)) / np.maximum(n, 1), 0.0)
        assert np.allclose(out[..., 2], conv(mean), rtol=2e-5, atol=1e-9)
    assert ths[0, 1, 2] == 0 and bsw[0, 1, 2] == np.float32(1.0) / np.float32(1e-8)   # G28: bsw = 1e8
    assert np.all(ths[..., [0, 1, 3, 4, 5, 6, 7]] == 0)   # only the requested layer is written


def test_oracle_regrid_equals_the_translated_reference():
    """INIT.f90:575-632 as translated from the reference's own source (oracle/_ref): bit for bit."""
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref not built")
    lib = ref_py.load()
    fields = fine_fields()
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    for layer in (1, 4, 8):
        got = [np.zeros((LAT_C, LON_C, 8), np.float32) for _ in range(4)]
        assert lib.h9r_regrid_soil_layer(LON_C, LAT_C, layer, *[p(a) for a in fields], *[p(a) for a in got]) == 0
        for a, b in zip(got, oracle_regrid(fields, layer)):
            assert np.array_equal(a, b)


@pytest.mark.gpu
def test_gpu_regrid_is_bit_exact():
    fields = fine_fields()
    h = H9(0)
    got = [np.zeros((LAT_C, LON_C, 8), np.float32) for _ in range(4)]
    for layer in (1, 8):
        h.regrid_soil_layer(LON_C, LAT_C, layer, *fields, *got)
        ref = oracle_regrid(fields, layer)
        for a, b in zip(got, ref):
            assert np.array_equal(a[..., layer - 1], b[..., layer - 1])
    assert h.counters()["launches"] >= 2
    h.close()
