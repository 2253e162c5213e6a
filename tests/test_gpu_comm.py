"""The multi-GPU entries of the C ABI (h9_comm_init / h9_annual_collective /
h9_get_gathered_annual / h9_get_budget): NCCL on the library's own stream (`-m gpu`).

World size 1 runs on any GPU box (NCCL accepts a one-rank communicator): it checks the
whole call sequence and that the gathered planes and the budget equal what h9_annual_device
and h9_get_annual report.  World size 2 needs two GPUs (skipped otherwise): two processes,
one ctx each, latitude bands of one grid; the id travels through a file, as MPI_Bcast would
carry it for the Fortran host; every rank must end with both bands' planes and the
all-reduced budget, and the scattered global field must equal a single-GPU run of the
whole grid (thread-per-cell kernel pinned: bit for bit)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch  # noqa: F401  (first: the process then holds torch's libnccl.so.2, which h9_comm_init
#                            reuses; loading the system copy first would clash with torch's import)

from helpers import THREAD_PER_CELL, make_gpu
from hybrid9_b200 import MATH_FAST, synth
from hybrid9_b200.host import comm_unique_id
from hybrid9_b200.state import init_state

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANES = ("npp", "plant_mass", "rnf", "evap", "theta_total")


def test_one_rank_collective_matches_the_local_views():
    w = synth.make_world(nx=72, ny=36, seed=9)
    nd = 5
    f = synth.make_forcing(w, nd, seed=9)
    h = make_gpu(w, mode=MATH_FAST, nyr=2)
    h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER))
    h.comm_init(1, 0, comm_unique_id())
    assert list(h.comm_land_counts()) == [h.num_land]
    for iy in (1, 2):
        assert h.run_days(np.full(nd, iy, np.int32), f) == 0
        h.annual_collective(iy)  # asynchronous on the ctx's stream
    ann = h.get_annual(2)
    g = h.get_gathered_annual(0, h.num_land)
    land = w.land
    for k, name in enumerate(PLANES):
        assert np.array_equal(g[k], ann[name][land]), name
    for i in range(8):
        assert np.array_equal(g[5 + i], ann["theta"][land][:, i])
    for iy in (1, 2):
        b = h.get_budget(iy)
        a = h.get_annual(iy)
        assert b[5] == h.num_land and b[7] == 0
        assert np.isclose(b[2], a["rnf"][land].astype(np.float64).sum(), rtol=1e-12)
        assert np.isclose(b[3], a["npp"][land].astype(np.float64).sum(), rtol=1e-12)
    st = h.get_state()
    b2 = h.get_budget(2)
    assert np.isclose(b2[0], st.h2osoi_liq[land].astype(np.float64).sum(), rtol=1e-12)
    assert np.isclose(b2[1], st.wa[land].astype(np.float64).sum(), rtol=1e-12)
    h.comm_destroy()
    h.close()


def test_collective_without_communicator_is_an_error():
    from hybrid9_b200.host import H9Error
    w = synth.make_world(nx=36, ny=18, seed=9)
    h = make_gpu(w, mode=MATH_FAST)
    h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER))
    with pytest.raises(H9Error):
        h.annual_collective(1)
    h.close()


RANK_SCRIPT = r"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from helpers import THREAD_PER_CELL, make_gpu
from hybrid9_b200 import MATH_FAST, synth, distributed as h9d
from hybrid9_b200.host import comm_unique_id
from hybrid9_b200.state import init_state
rank, nranks, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
w = synth.make_world(nx=144, ny=72, seed=5)
nd = 6
f = synth.make_forcing(w, nd, seed=3)
sub, lat_s, lat_c, n_land = h9d.shard_world(w, rank, nranks)
fs = h9d.shard_forcing(f, lat_s, lat_c)
h = make_gpu(sub, mode=MATH_FAST, device=rank, block=THREAD_PER_CELL)
h.set_state(init_state(sub.soil_tex, sub.theta_s, synth.ZI_DRIVER))
if rank == 0:
    with open(idfile + ".tmp", "wb") as fh: fh.write(comm_unique_id())
    os.replace(idfile + ".tmp", idfile)          # what MPI_Bcast does for the Fortran host
else:
    while not os.path.exists(idfile): time.sleep(0.05)
uid = open(idfile, "rb").read()
h.comm_init(nranks, rank, uid)
counts = [int(x) for x in h.comm_land_counts()]
assert counts == [int(x) for x in n_land], (counts, n_land)
assert h.run_days(np.ones(nd, np.int32), fs) == 0
h.annual_collective(1)
parts, budget = h9d.fetch_gathered(h, 1, counts)
np.savez(out, budget=budget, lat_s=lat_s, land_index=h.land_index(), **{{f"part{{r}}": p for r, p in enumerate(parts)}})
h.comm_destroy(); h.close()
"""


def test_two_ranks_gather_and_budget():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    nranks = 2
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "rank.py")
        with open(script, "w") as fh:
            fh.write(RANK_SCRIPT.format(root=ROOT))
        idfile = os.path.join(td, "nccl.id")
        outs = [os.path.join(td, f"out{r}.npz") for r in range(nranks)]
        procs = [subprocess.Popen([sys.executable, script, str(r), str(nranks), idfile, outs[r]])
                 for r in range(nranks)]
        for p in procs:
            assert p.wait(timeout=600) == 0
        res = [np.load(o) for o in outs]
    # every rank holds the same gathered planes and the same budget
    for r in range(nranks):
        assert np.array_equal(res[0][f"part{r}"], res[1][f"part{r}"])
    assert np.array_equal(res[0]["budget"], res[1]["budget"])
    # against one GPU stepping the whole grid
    from hybrid9_b200 import distributed as h9d
    w = synth.make_world(nx=144, ny=72, seed=5)
    f = synth.make_forcing(w, 6, seed=3)
    h = make_gpu(w, mode=MATH_FAST, block=THREAD_PER_CELL)
    h.set_state(init_state(w.soil_tex, w.theta_s, synth.ZI_DRIVER))
    assert h.run_days(np.ones(6, np.int32), f) == 0
    ann = h.get_annual(1)
    h.close()
    grid = h9d.scatter_to_grid([res[0][f"part{r}"] for r in range(nranks)],
                               [res[r]["land_index"] for r in range(nranks)],
                               [int(res[r]["lat_s"]) for r in range(nranks)], w.nx, w.ny)
    for k, name in enumerate(PLANES):
        assert np.array_equal(grid[k], ann[name], equal_nan=True), name
    assert res[0]["budget"][5] == int(w.land.sum())
    assert np.isclose(res[0]["budget"][2], ann["rnf"][w.land].astype(np.float64).sum(), rtol=1e-9)
