"""Calendar of the reference driver (host side, integer arithmetic only).

time_BOY: INIT.f90:844-859 -- day number (1 = 1 Jan 1860) of 1 January of each
year 1860..2300, with the reference's own leap rule (it tests ``jyear-1``, i.e.
the year that has just ended).  The decade/year loop bounds follow
HYBRID9.f90:103-113,130,150,156,269.
"""
from __future__ import annotations

import numpy as np


def time_boy(year: int) -> int:
    if not 1860 <= year <= 2300:
        raise ValueError("time_BOY is defined for 1860..2300 (INIT.f90:844-859)")
    t = 1
    for jyear in range(1861, year + 1):
        prev = jyear - 1
        if prev % 4 != 0:
            t += 365
        elif prev % 100 != 0:
            t += 366
        elif prev % 400 != 0:
            t += 365
        else:
            t += 366
    return t


def decade_years(idec: int) -> tuple[int, int]:
    """(syr, eyr) of decade iDEC (1 = 1901-1910; 12 = 2011-2012), HYBRID9.f90:103-113."""
    syr = (idec - 1) * 10 + 1901
    eyr = syr + 9 if idec < 12 else syr + 1
    return syr, eyr


def decade_days(idec: int) -> int:
    """NTIMES of the decade's PGF files: days from 1 Jan syr to 31 Dec eyr."""
    syr, eyr = decade_years(idec)
    return time_boy(eyr + 1) - time_boy(syr)


def year_index_of_days(idec: int, idec_start: int = 1) -> np.ndarray:
    """iY (HYBRID9.f90:269, 1-based year within the run) for every day iT of decade iDEC."""
    syr, eyr = decade_years(idec)
    out = []
    for jyear in range(syr, eyr + 1):
        n = time_boy(jyear + 1) - time_boy(jyear)
        iy = jyear - ((idec_start - 1) * 10 + 1901) + 1
        out.append(np.full(n, iy, dtype=np.int32))
    return np.concatenate(out)
