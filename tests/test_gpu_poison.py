"""No kernel may read device memory that nothing wrote (`-m gpu`).

compute-sanitizer (initcheck) is not available on this GPU pool (profiles/r02/
sanitizer_unavailable.txt); this is the stand-in.  With H9_POISON=1 every device array that the
library's contract says is completely written before it is read -- state, parameters, the
per-call forcing and diagnostic buffers -- starts as 0xFF bytes (NaN as float, -1 as int) instead
of zeros.  The whole call sequence must then give the SAME BITS as the normal run, for every
kernel variant: the reference leaves `smp` uninitialised (INIT.f90:109, read at HYDROLOGY.f90:271
before :633 writes it), which is exactly the kind of read this would expose."""
import os

import numpy as np
import pytest

from helpers import FAST_KERNEL_IDS, FAST_KERNELS, assert_state_equal, make_gpu
from hybrid9_b200 import MATH_EXACT, MATH_FAST, synth
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu


def scenario(w, f, mode, block, poison):
    old = os.environ.get("H9_POISON")
    os.environ["H9_POISON"] = "1" if poison else "0"
    try:
        h = make_gpu(w, mode=mode, nyr=2, block=block)
        st = synth.randomize_state(w, init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER), seed=11)
        h.set_state(st, with_smp=False)  # smp: the library's defined zero, not device garbage
        out = [h.hydrology_step({k: np.ascontiguousarray(v[0]) for k, v in f.items()})]
        out.append(h.grow_day(np.ascontiguousarray(f["tas"][0])))
        rc = h.run_days(np.array([1, 1, 2], np.int32), f)
        p, ds, ps = h.pack_forcing(f, 3)
        rc |= h.run_days_device(np.array([2, 2, 2], np.int32), p, ds, ps)
        res = (rc, h.get_state(), h.get_annual(1), h.get_annual(2), out, h.get_fault())
        h.close()
        return res
    finally:
        if old is None:
            os.environ.pop("H9_POISON", None)
        else:
            os.environ["H9_POISON"] = old


@pytest.mark.parametrize("mode,block", [(MATH_EXACT, 0)] + [(MATH_FAST, b) for b in FAST_KERNELS],
                         ids=["exact"] + FAST_KERNEL_IDS)
def test_poisoned_allocations_change_nothing(mode, block):
    w = synth.make_world(nx=72, ny=36, seed=9)
    f = synth.make_forcing(w, 3, seed=9)
    a = scenario(w, f, mode, block, poison=False)
    b = scenario(w, f, mode, block, poison=True)
    land = w.land
    assert a[0] == b[0] and a[5].n_faulted == b[5].n_faulted
    assert_state_equal(a[1], b[1], land)
    for x, y in ((a[2], b[2]), (a[3], b[3])):
        for k in x:
            assert np.array_equal(x[k], y[k], equal_nan=True), k
    for x, y in zip(a[4], b[4]):
        for k in x:
            if isinstance(x[k], np.ndarray):
                assert np.array_equal(x[k][land], y[k][land], equal_nan=True), k
    # and the poisoned run itself is NaN-free on land wherever the normal run is
    for n in ("h2osoi_liq", "zwt", "wa", "lai", "plant_mass", "smp", "rootr_col"):
        assert np.array_equal(np.isnan(getattr(a[1], n)[land]), np.isnan(getattr(b[1], n)[land])), n
