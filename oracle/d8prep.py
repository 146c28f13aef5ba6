"""CPU reference for the device D8 preprocessing (wfa_flowdir_from_dem) -- TEST INFRASTRUCTURE ONLY.

The reference shells out to TauDEM for this once-per-call step: ``pitremove -z dem.tif -fel CONDEM.tif``
and ``D8Flowdir -p FlowDir.tif -fel CONDEM.tif`` (R/Routing_GR2MSemiDistr.R:88-93).  TauDEM is not
available and its flat-resolution tie-breaking (Garbrecht & Martz) is not reproduced; the rule below
is this package's own, chosen to be order-independent so that a parallel device implementation
can match it bit for bit:

  1. outlets = valid cells with an invalid / off-grid neighbour; filled[c] = z[c] there, elsewhere the
     lowest elevation to which c must be raised to drain to an outlet (the unique depression fill);
  2. a cell with a strictly lower neighbour on the filled surface points to the steepest one (drop /
     distance, sqrt(2) on diagonals, first k = 1..8 on ties; TauDEM codes 1=E .. 8=SE);
  3. an outlet without a lower neighbour points at its first invalid / off-grid neighbour;
  4. the remaining cells lie on flats: breadth-first distance, through cells of equal filled elevation,
     to the nearest cell resolved by 2 or 3; a flat cell points at its first neighbour of equal
     elevation that is one step closer.
"""
from __future__ import annotations

import heapq

import numpy as np

DX = (0, 1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, 0, -1, -1, -1, 0, 1, 1, 1)
NODATA = -32768
SQRT2 = 1.4142135623730951


def valid_mask(dem, nodata=None):
    z = np.asarray(dem, np.float64)
    v = np.isfinite(z) & (z > -1e30)
    if nodata is not None:
        v &= ~np.isclose(z, nodata, rtol=1e-6)
    return v


def _shift(a, k, fill):
    nrow, ncol = a.shape
    p = np.pad(a, 1, constant_values=fill)
    return p[1 + DY[k]:1 + DY[k] + nrow, 1 + DX[k]:1 + DX[k] + ncol]


def fill_and_flowdir(dem, nodata=None):
    """dem [nrow, ncol]; returns (filled float64 with NaN at invalid cells, flowdir int16)."""
    z = np.asarray(dem, np.float64)
    nrow, ncol = z.shape
    valid = valid_mask(z, nodata)
    edge = np.zeros((nrow, ncol), bool)
    for k in range(1, 9):
        edge |= ~_shift(valid, k, False)
    edge &= valid
    # 1. unique depression fill by priority flood (any correct algorithm gives the same surface)
    filled = np.where(valid, z, np.nan)
    done = ~valid
    heap = []
    for r, c in zip(*np.nonzero(edge)):
        heapq.heappush(heap, (filled[r, c], int(r) * ncol + int(c)))
        done[r, c] = True
    while heap:
        zc, i = heapq.heappop(heap)
        r, c = divmod(i, ncol)
        for k in range(1, 9):
            r2, c2 = r + DY[k], c + DX[k]
            if 0 <= r2 < nrow and 0 <= c2 < ncol and not done[r2, c2]:
                done[r2, c2] = True
                if filled[r2, c2] < zc:
                    filled[r2, c2] = zc
                heapq.heappush(heap, (filled[r2, c2], r2 * ncol + c2))
    # 2./3. steepest descent, outlets draining out of the domain
    best = np.zeros((nrow, ncol))
    code = np.zeros((nrow, ncol), np.int16)
    first_out = np.zeros((nrow, ncol), np.int16)
    for k in range(1, 9):
        nb = _shift(filled, k, np.nan)
        with np.errstate(invalid="ignore"):
            s = (filled - nb) / (1.0 if k % 2 == 1 else SQRT2)
            better = s > best
        best = np.where(better, s, best)
        code = np.where(better, k, code).astype(np.int16)
        first_out = np.where((first_out == 0) & np.isnan(nb), k, first_out).astype(np.int16)
    code = np.where((code == 0) & edge, first_out, code)
    code = np.where(valid, code, 0).astype(np.int16)
    # 4. flats: level-synchronous BFS through equal filled elevation
    dist = np.where(valid & (code > 0), 0, -1).astype(np.int32)
    t = 0
    while True:
        t += 1
        todo = valid & (dist < 0)
        if not todo.any():
            break
        newcode = np.zeros((nrow, ncol), np.int16)
        for k in range(8, 0, -1):                      # descending so that the first k wins
            nb_f = _shift(filled, k, np.nan)
            nb_d = _shift(dist, k, -1)
            hit = todo & (nb_f == filled) & (nb_d == t - 1)
            newcode = np.where(hit, k, newcode).astype(np.int16)
        got = newcode > 0
        if not got.any():
            break                                      # unreachable flats (cannot happen on a filled surface)
        code = np.where(got, newcode, code).astype(np.int16)
        dist = np.where(got, t, dist).astype(np.int32)
    code = np.where(valid & (code > 0), code, NODATA).astype(np.int16)
    return filled, code
