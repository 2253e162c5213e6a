"""Flat restart file of hybrid9_b200/spinup.py (N3): round trip, layout and error handling (CPU)."""
import numpy as np
import pytest

from hybrid9_b200 import load_restart, save_restart, synth
from hybrid9_b200.state import init_state


def test_round_trip_and_layout(tmp_path):
    w = synth.make_world(nx=24, ny=12, seed=17)
    st = synth.randomize_state(w, init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), seed=3)
    p = str(tmp_path / "r.h9")
    save_restart(p, st)
    back = load_restart(p)
    for n in st.names():
        assert np.array_equal(getattr(back, n), getattr(st, n)) and getattr(back, n).dtype == getattr(st, n).dtype, n
    raw = open(p, "rb").read()
    assert len(raw) == 64 + 4 * sum(getattr(st, n).size for n in st.names())
    # first array after the header is h2osoi_liq in the reference's (8,lon_c,lat_c) memory order
    first = np.frombuffer(raw[64:64 + 4 * st.h2osoi_liq.size], "<f4")
    assert np.array_equal(first, st.h2osoi_liq.ravel())


def test_bad_files_are_refused(tmp_path):
    w = synth.make_world(nx=24, ny=12, seed=17)
    st = init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER)
    p = str(tmp_path / "r.h9")
    save_restart(p, st)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-8])
    with pytest.raises(ValueError):
        load_restart(p)
    open(p, "wb").write(b"X" + raw[1:])
    with pytest.raises(ValueError):
        load_restart(p)
    open(p, "wb").write(raw + b"\0")
    with pytest.raises(ValueError):
        load_restart(p)
