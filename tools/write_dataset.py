"""Write a synthetic data set in the flat-file layout hybrid9_b200/host_cpp/h9_driver reads:
the arrays a Fortran host would hold after INIT and READ_PGF, in the reference's memory
order, plus a driver.txt in the reference's own positional format (EXECUTE/driver.txt)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid9_b200 import calendar, synth  # noqa: E402

DRIVER_TXT = """'{out}' ! Path for output
{nisurf}  ! NISURF     (surface timesteps per day; 48 == 1800 s/dt)
.T. ! Use PGF forcing (PGF)?
{d0:2d}  ! iDEC_start ( 1 = 1901-1910)
{d1:2d}  ! iDEC_end   (11 = 2001-2010; 12 = 2011-2012)
.F. ! INTERACTIVE (following options only operate if .TRUE.)
  .F.         ! Local climate (LCLIM)?
  'LCLIM/climate_US-Var_new.csv' ! Local clm filename  (LCLIM_filename)
  'LCLIM/US-Var_soil.csv'        ! Local soil filename (LSOIL_filename)
  2002                          ! First year in local climate file (syr)
  2003                          ! Last year in local climate file (eyr)
    10                          ! No. years to spin-up (NYR_SPIN_UP)
-120.95       ! Local longitude (lon_w)
  38.41       ! Local latitude  (lat_w)
   1          ! lon_c_w (count towards east)
   1          ! lat_c_w (count towards south)
{zi}
"""


def write_dataset(path: str, world, idec_start: int, idec_end: int, nisurf: int = 48, seed: int = 9):
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "grid.txt"), "w") as f:
        f.write(f"{world.nx} {world.ny}\n")
    world.soil_tex.astype("<i4").tofile(os.path.join(path, "soil_tex.i32"))
    for k in ("theta_s", "hksat", "bsw", "psi_s", "fmax"):
        getattr(world, k).astype("<f4").tofile(os.path.join(path, f"{k}.f32"))
    zi = "\n".join(f"{v:8.1f}" + (" ! Soil interface depth (zi (0))." if i == 0 else "")
                   for i, v in enumerate(synth.ZI_DRIVER))
    with open(os.path.join(path, "driver.txt"), "w") as f:
        f.write(DRIVER_TXT.format(out=path, nisurf=nisurf, d0=idec_start, d1=idec_end, zi=zi))
    forcing = {}
    for idec in range(idec_start, idec_end + 1):
        nd = calendar.decade_days(idec)
        fo = synth.make_forcing(world, nd, seed=seed + idec)
        for k, v in fo.items():
            v.astype("<f4").tofile(os.path.join(path, f"{k}_dec{idec:02d}.f32"))
        forcing[idec] = fo
    return forcing


if __name__ == "__main__":
    out = sys.argv[1]
    nx, ny = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (72, 36)
    d0, d1 = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (12, 12)
    write_dataset(out, synth.make_world(nx=nx, ny=ny, seed=9), d0, d1)
    print("wrote", out)
