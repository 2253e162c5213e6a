"""The drop-in boundary: libh9gpu.so loads without a GPU, exports every symbol that
include/h9gpu.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from hybrid9_b200 import host, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "h9gpu.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h9_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = host.load_library()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/h9gpu.h but not exported"
    # and the Python host binds exactly the declared set
    assert sorted(host.ABI) == names


def test_exported_symbols_are_plain_c_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", host.library_path()], capture_output=True,
                         text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in header_symbols():
        assert n in exported  # unmangled: extern "C"


def test_product_library_does_not_link_the_oracle_or_torch():
    out = subprocess.run(["ldd", host.library_path()], capture_output=True, text=True).stdout
    assert "h9oracle" not in out and "torch" not in out and "python" not in out
    syms = subprocess.run(["nm", "-D", host.library_path()], capture_output=True, text=True).stdout
    assert "h9o_" not in syms and "h9t_" not in syms


def test_no_cpu_fallback(gpu_available):
    """Without a CUDA device h9_create must fail loudly; the host layer raises."""
    if gpu_available:
        pytest.skip("a GPU is present: the failure path is exercised on the CPU-only box")
    lib = host.load_library()
    h = C.c_void_p()
    assert lib.h9_create(C.byref(h), -1) == -2 and not h  # H9_ERR_CUDA
    with pytest.raises(host.H9Error):
        host.H9()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("H9GPU_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(host, "_LIB", None)
    with pytest.raises(host.H9Error, match="no CPU fallback"):
        host.load_library()


def test_null_context_is_rejected_not_dereferenced():
    lib = host.load_library()
    assert lib.h9_configure(None, 1, 1, 48, None, 1) == -1
    assert lib.h9_num_land(None) == -1
    assert lib.h9_destroy(None) == -1
    assert lib.h9_last_error(None) == b"null ctx"


def test_partition_lat_bands_is_a_pure_host_function():
    w = synth.make_world(nx=72, ny=36, seed=9)
    for nranks in (1, 2, 3, 8):
        lat_s, lat_c, n_land = host.partition_lat_bands(w.soil_tex, w.theta_s, nranks)
        assert lat_s[0] == 1 and lat_c.sum() == w.ny and n_land.sum() == w.land.sum()
        assert np.array_equal(lat_s[1:], (lat_s + lat_c)[:-1])       # contiguous, in order
        rows = w.land.sum(axis=1)
        for r in range(nranks):
            assert n_land[r] == rows[lat_s[r] - 1: lat_s[r] - 1 + lat_c[r]].sum()
        if nranks > 1:
            assert n_land.max() - n_land.min() <= 2 * rows.max()     # balanced by land count
