"""Compare the outputs of oracle/_ref/h9_ref_driver (the reference's own Fortran HYDROLOGY +
GROW, when a Fortran compiler exists) with those of hybrid9_b200/host_cpp/h9_driver (GPU) in
the same data directory.  Prints per-field error statistics; exit code 1 if the soil-water
p99.9 relative error exceeds the exact-mode gate of tests/test_gpu_parity.py."""
import sys

import numpy as np

d = sys.argv[1]
nx, ny = (int(v) for v in open(f"{d}/grid.txt").read().split())
soil = np.fromfile(f"{d}/soil_tex.i32", "<i4").reshape(ny, nx)
ths = np.fromfile(f"{d}/theta_s.f32", "<f4").reshape(ny, nx, 8)
land = (soil > 0) & (soil != 13) & (ths.sum(axis=2) > 1e-8)
bad = False
for name, per in (("state_h2osoi_liq", 8), ("state_zwt", 1), ("state_plant_mass", 1), ("axy_rnf", 1),
                  ("axy_npp", 1), ("axy_theta", 8)):
    a = np.fromfile(f"{d}/out_{name}.f32", "<f4")
    b = np.fromfile(f"{d}/out_ref_{name}.f32", "<f4")
    a = a.reshape(-1, ny, nx, per) if per > 1 else a.reshape(-1, ny, nx)
    b = b.reshape(a.shape)
    x, y = a[:, land].astype(np.float64), b[:, land].astype(np.float64)
    rel = np.abs(x - y) / np.maximum(np.abs(y), 1e-3)
    print(f"{name:20s} p50 {np.median(rel):.3e}  p99.9 {np.quantile(rel, 0.999):.3e}  max {rel.max():.3e}")
    if name == "state_h2osoi_liq" and np.quantile(rel, 0.999) > 5e-3:
        bad = True
sys.exit(1 if bad else 0)
