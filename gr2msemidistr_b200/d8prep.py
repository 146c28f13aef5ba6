"""Host-side replacement for the two TauDEM preprocessing calls of Routing_GR2MSemiDistr:
``pitremove -z dem.tif -fel CONDEM.tif`` and ``D8Flowdir -p FlowDir.tif -fel CONDEM.tif``
(R/Routing_GR2MSemiDistr.R:88-93).  They run once per call, are not on the per-month path, and
TauDEM is not available here, so the rule below is this package's own and is documented as such:

  1. Priority-flood fill (Barnes et al. 2014): cells on the grid edge or next to nodata are outlets;
     every other cell is raised to the lowest spill elevation that reaches an outlet.
  2. Direction: steepest descent on the filled surface over the 8 neighbours (drop / distance,
     distance sqrt(2) on diagonals, first k = 1..8 on ties, TauDEM's code order 1=E .. 8=SE).  Cells
     with no strictly lower neighbour (flats, filled pits) point at the neighbour they were flooded
     from, which always leads to an outlet without cycles.  Outlet cells with no lower neighbour point
     at their first nodata / off-grid neighbour, i.e. they drain out of the domain.

TauDEM resolves flats with Garbrecht & Martz; where the DEM has flats the two grids will differ.
That deviation is confined to this preprocessing step -- the routing kernels take any D8 grid.
"""
from __future__ import annotations

import heapq

import numpy as np

DX = (0, 1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, 0, -1, -1, -1, 0, 1, 1, 1)
NODATA = -32768


def fill_and_flowdir(dem, nodata=None):
    """dem [nrow, ncol] float; returns (filled float64 [nrow, ncol], flowdir int16 TauDEM codes)."""
    z = np.asarray(dem, np.float64)
    nrow, ncol = z.shape
    valid = np.isfinite(z)
    if nodata is not None:
        valid &= ~np.isclose(z, nodata, rtol=1e-6) & (z > -1e30)
    filled = np.where(valid, z, np.nan)
    parent = np.zeros((nrow, ncol), np.int8)          # code of the neighbour a cell was flooded from
    done = ~valid
    heap = []
    vp = np.pad(valid, 1, constant_values=False)
    edge = np.zeros((nrow, ncol), bool)
    for k in range(1, 9):
        edge |= ~vp[1 + DY[k]:1 + DY[k] + nrow, 1 + DX[k]:1 + DX[k] + ncol]
    edge &= valid
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
                parent[r2, c2] = ((k - 1 + 4) % 8) + 1   # from (r2,c2) back towards (r,c)
                heapq.heappush(heap, (filled[r2, c2], r2 * ncol + c2))
    fp = np.pad(filled, 1, constant_values=np.nan)
    best = np.zeros((nrow, ncol))
    code = np.zeros((nrow, ncol), np.int16)
    first_out = np.zeros((nrow, ncol), np.int16)
    for k in range(1, 9):
        nb = fp[1 + DY[k]:1 + DY[k] + nrow, 1 + DX[k]:1 + DX[k] + ncol]
        with np.errstate(invalid="ignore"):
            s = (filled - nb) / (1.0 if k % 2 == 1 else np.sqrt(2.0))
            better = s > best
        best = np.where(better, s, best)
        code = np.where(better, k, code).astype(np.int16)
        first_out = np.where((first_out == 0) & np.isnan(nb), k, first_out).astype(np.int16)
    code = np.where(code == 0, parent.astype(np.int16), code)
    code = np.where((code == 0) & edge, first_out, code)
    code = np.where(valid & (code > 0), code, NODATA).astype(np.int16)
    return filled, code
