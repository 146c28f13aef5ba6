"""R >= 4.0 ``round(x, digits)`` for the host-side wrappers.

The reference rounds what it returns (R/Run_GR2MSemiDistr.R:231-236,255; R/Routing_GR2MSemiDistr.R:131;
R/Optim_GR2MSemiDistr.R:232-236).  The C ABI hands back unrounded doubles; the wrappers apply R's
algorithm here: candidates floor/ceil(x*10^d)/10^d, the closer one in double arithmetic wins, the
even last digit on a tie -- which is not ``numpy.round`` (0.15 -> 0.1, 2.5 -> 2, 29.65 -> 29.6).
"""
from __future__ import annotations

import numpy as np


def rround(x, digits: int = 0):
    x = np.asarray(x, dtype=np.float64)
    if digits == 0:
        return np.rint(x)
    a = np.abs(x)
    p10 = np.float64(10.0) ** abs(int(digits))
    if digits > 0:
        x10 = p10 * a
        i10 = np.floor(x10)
        xd, xu = i10 / p10, np.ceil(x10) / p10
    else:
        x10 = a / p10
        i10 = np.floor(x10)
        xd, xu = i10 * p10, np.ceil(x10) * p10
    du, dd = xu - a, a - xd
    r = np.where((dd < du) | ((dd == du) & (np.mod(i10, 2.0) == 0.0)), xd, xu)
    with np.errstate(divide="ignore", invalid="ignore"):
        e10 = np.floor(np.log10(a)) + 1
    keep = ~np.isfinite(x) | (x == 0) | (digits + e10 > 15)
    return np.where(keep, x, np.copysign(r, x))
