"""Generate tests/golden/h9_golden_v1.npz: a frozen input/output pair of the oracle.

The reference ships no golden vectors and cannot be run here, so this
fixture is NOT a reference output: it freezes what oracle/h9_oracle.cpp (strict build)
produces for a small seeded case, so that (a) an accidental change of the oracle is caught
on CPU, and (b) the GPU can be compared with a committed vector without running the oracle.
Regenerate (only after a deliberate oracle change): python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import STATE_FIELDS, make_oracle  # noqa: E402
from hybrid9_b200 import synth  # noqa: E402

NX, NY, SEED, NDAYS, NISURF = 24, 12, 17, 12, 48


def build():
    w = synth.make_world(nx=NX, ny=NY, seed=SEED, n_class13=2, n_zero_theta=2)
    f = synth.make_forcing(w, NDAYS, seed=SEED)
    out = {}
    for tag in ("init", "random"):
        o = make_oracle(w, nisurf=NISURF, nyr=2)
        o.init_state()
        st0 = o.get_state() if tag == "init" else synth.randomize_state(w, o.get_state(), seed=SEED)
        o.set_state(st0)
        yi = np.concatenate([np.full(7, 1, np.int32), np.full(NDAYS - 7, 2, np.int32)])
        rc = o.run_days(yi, f)
        st1 = o.get_state()
        out[f"{tag}_rc"] = np.int32(rc)
        for n in STATE_FIELDS + ("nplants",):
            out[f"{tag}_in_{n}"] = getattr(st0, n)
            out[f"{tag}_out_{n}"] = getattr(st1, n)
        for iy in (1, 2):
            for k, v in o.get_annual(iy).items():
                out[f"{tag}_axy{iy}_{k}"] = v
    out["year_index"] = yi
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "h9_golden_v1.npz"), **build())
    print("wrote", os.path.join(HERE, "h9_golden_v1.npz"), os.path.getsize(os.path.join(HERE, "h9_golden_v1.npz")), "bytes")
