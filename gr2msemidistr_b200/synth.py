"""Seeded synthetic inputs of the BASELINE.json shapes: forcing tables in the layout
Create_Forcing_Inputs produces (R/Create_Forcing_Inputs.R:185-189), candidate parameter sets inside
the documented calibration bounds (R/Optim_GR2MSemiDistr.R:29-31), observed discharge with NA gaps
(raw/qobs.txt has 12), and pit-free D8 grids in TauDEM codes (1=E .. 8=SE, counter-clockwise).

The same arrays feed the CUDA path and the CPU oracle in tests/ and bench.py.
"""
from __future__ import annotations

import numpy as np

SEED = 20261019
SET0 = (10.976, 0.665, 1.186, 1.169)          # documented example set, R/Run_GR2MSemiDistr.R:33
BOUNDS_MIN = (1.0, 0.01, 0.8, 0.8)            # R/Optim_GR2MSemiDistr.R:30
BOUNDS_MAX = (2000.0, 2.0, 1.2, 1.2)          # R/Optim_GR2MSemiDistr.R:31

_DAYS = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)


def days_in_month(year: int, month: int) -> int:
    """numberOfDays(), R/Run_GR2MSemiDistr.R:108-117."""
    if month == 2 and ((year % 4 == 0 and year % 100 != 0) or year % 400 == 0):
        return 29
    return _DAYS[month - 1]


def ndays(ntime: int, start_year: int = 1981, start_month: int = 1) -> np.ndarray:
    out = np.empty(ntime, np.int32)
    y, m = start_year, start_month
    for t in range(ntime):
        out[t] = days_in_month(y, m)
        m += 1
        if m > 12:
            y, m = y + 1, 1
    return out


def forcing(ncell: int, ntime: int, seed: int = SEED, dtype=np.float64):
    """P, E as [ncell, ntime] C-contiguous (the bytes of R's column-major [ntime x ncell] table),
    1-decimal values like Create_Forcing_Inputs' round(,1) (Create:114); area [km2]; ndays."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(ntime)
    a = rng.lognormal(0.0, 0.5, ncell).astype(np.float32)
    phi = rng.uniform(0, 2 * np.pi, ncell).astype(np.float32)
    P = np.empty((ncell, ntime), dtype)
    E = np.empty((ncell, ntime), dtype)
    step = max(1, (1 << 22) // max(ntime, 1))
    season = (2 * np.pi * (t % 12) / 12).astype(np.float32)
    for c0 in range(0, ncell, step):
        c1 = min(ncell, c0 + step)
        theta = 65.0 * (1.0 + 0.8 * np.sin(season[None, :] + phi[c0:c1, None]))
        g = rng.standard_gamma(0.9, (c1 - c0, ntime)).astype(np.float32) * theta * a[c0:c1, None]
        P[c0:c1] = np.round(g.astype(np.float64), 1)
        e = 122.0 + 40.0 * np.sin(season)[None, :] + rng.normal(0, 10, (c1 - c0, ntime)).astype(np.float32)
        E[c0:c1] = np.round(np.clip(e, 0, 210).astype(np.float64), 1)
    area = rng.lognormal(np.log(400.0), 0.6, ncell)
    return P, E, area, ndays(ntime)


def param_sets(nsets: int, nreg: int = 1, seed: int = SEED + 1) -> np.ndarray:
    """[nsets, 4*nreg] in the reference's order [X1 x nreg | X2 x nreg | Fpp x nreg | Fpe x nreg]
    (R/Run_GR2MSemiDistr.R:85-88); set 0 is the documented example set."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((nsets, 4 * nreg))
    for j in range(4):
        out[:, j * nreg:(j + 1) * nreg] = rng.uniform(BOUNDS_MIN[j], BOUNDS_MAX[j], (nsets, nreg))
        out[0, j * nreg:(j + 1) * nreg] = SET0[j]
    return out


def qobs_from(sim: np.ndarray, seed: int = SEED + 2, na_frac: float = 0.03) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    q = np.round(np.asarray(sim, np.float64) * rng.lognormal(0.0, 0.2, len(sim)), 2)
    q[rng.random(len(sim)) < na_frac] = np.nan
    return q


def d8_grid(nrow: int, ncol: int, seed: int = SEED + 3, device="cpu", nodata_frac: float = 0.02):
    """Pit-free synthetic D8 grid as an int16 torch tensor [nrow, ncol] on `device`.

    Elevation = plane tilted to the south-east + smooth large-scale relief + white roughness whose
    cell-to-cell amplitude stays below the tilt, so every interior cell has a strictly lower
    neighbour (no pit filling needed).  Direction = steepest descent over the 8 neighbours with
    diagonal distance sqrt(2), first k on ties; cells on the outer ring and a `nodata_frac` share of 64 x 64
    blocks are nodata (-32768); a cell whose steepest neighbour is nodata still points at it (it
    drains out of the domain there)."""
    import torch
    import torch.nn.functional as F

    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = torch.device(device)
    rr = torch.arange(nrow, device=dev, dtype=torch.float32)[:, None]
    cc = torch.arange(ncol, device=dev, dtype=torch.float32)[None, :]
    z = -(0.55 * rr + 0.45 * cc)
    for k, amp in ((64, 3.0), (16, 0.75)):
        hr, hc = max(2, nrow // k + 2), max(2, ncol // k + 2)
        coarse = torch.rand((1, 1, hr, hc), generator=g).to(dev)
        z = z + amp * F.interpolate(coarse, size=(nrow, ncol), mode="bilinear", align_corners=True)[0, 0]
    # white roughness from a counter hash so it can be made on the device without a CPU tensor
    idx = (rr.to(torch.int64) * ncol + cc.to(torch.int64)) + (seed & 0xFFFF) * 7919
    h = (idx * 2654435761) & 0xFFFFFFFF
    h = ((h ^ (h >> 15)) * 2246822519) & 0xFFFFFFFF
    h = ((h ^ (h >> 13)) * 3266489917) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    z = z + 0.8 * (h.to(torch.float32) / 4294967296.0)
    del idx
    del h
    # nodata "lakes": whole 64 x 64 blocks, so flow paths run for thousands of cells between sinks
    hb = (rr.to(torch.int64) // 64) * 40503 + (cc.to(torch.int64) // 64) * 9973 + (seed & 0xFFFF)
    hb = (hb * 2654435761) & 0xFFFFFFFF
    hb = ((hb ^ (hb >> 15)) * 2246822519) & 0xFFFFFFFF
    hb = (hb ^ (hb >> 13)) & 0xFFFFFFFF
    hole = hb.to(torch.float32) / 4294967296.0 < nodata_frac
    del hb
    hole[0, :] = True
    hole[-1, :] = True
    hole[:, 0] = True
    hole[:, -1] = True
    dx = (1, 1, 0, -1, -1, -1, 0, 1)
    dy = (0, -1, -1, -1, 0, 1, 1, 1)
    zp = F.pad(z[None, None], (1, 1, 1, 1), value=float("inf"))[0, 0]
    best = torch.zeros_like(z)
    code = torch.zeros((nrow, ncol), dtype=torch.int16, device=dev)
    for k in range(8):
        nb = zp[1 + dy[k]:1 + dy[k] + nrow, 1 + dx[k]:1 + dx[k] + ncol]
        s = (z - nb) / (1.0 if k % 2 == 0 else 1.4142135623730951)
        better = s > best
        best = torch.where(better, s, best)
        code = torch.where(better, torch.full_like(code, k + 1), code)
    code = torch.where(hole | (code == 0), torch.full_like(code, -32768), code)
    return code


def d8_convergent(n: int, device="cpu"):
    """A single convergent drainage tree on an n x n grid, by construction: cells above the main diagonal flow south,
    cells below it flow east, the diagonal flows south-east and leaves through the south-east corner (outer ring
    nodata; the row and column next to the north and west ring drain outwards, so the streams start on cells that
    TauDEM's edge-contamination rule leaves clean).  With block_polygons every block centre drains through the last
    block, whose interior the main stem crosses: one outlet polygon with the whole basin upstream, like a gauged
    watershed."""
    import torch
    dev = torch.device(device)
    rr = torch.arange(n, device=dev)[:, None]
    cc = torch.arange(n, device=dev)[None, :]
    code = torch.where(rr < cc, 7, torch.where(rr > cc, 1, 8)).to(torch.int16)
    code[1, :] = 3
    code[:, 1] = 5
    code[0, :] = -32768
    code[-1, :] = -32768
    code[:, 0] = -32768
    code[:, -1] = -32768
    return code


def fractal_dem(nrow: int, ncol: int, seed: int = SEED + 4, device="cpu", nodata_frac: float = 0.02, octaves: int = 5,
                relief: float = 1.0):
    """SURVEY 8(d) terrain: plane tilted to the south-east + `octaves` octaves of value noise whose local slopes
    are comparable to the tilt (so the surface is full of closed depressions and the drainage meanders), 2 % nodata in
    64 x 64 blocks plus the outer ring.  Returns (z float64 [nrow, ncol], valid bool) torch tensors on `device`;
    wfa_flowdir_from_dem fills it and derives the D8 grid."""
    import torch
    import torch.nn.functional as F

    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = torch.device(device)
    rr = torch.arange(nrow, device=dev, dtype=torch.float32)[:, None]
    cc = torch.arange(ncol, device=dev, dtype=torch.float32)[None, :]
    z = -(0.55 * rr + 0.45 * cc)
    # wavelengths 1024, 256, 64, 16, 4 cells; amplitude relief x tilt x wavelength: an octave's slope reaches relief x the tilt
    # where neighbouring lattice values differ by 1 (typical: a third of that); relief = 1 leaves ~4 % of a 20 000^2 grid in lakes
    for o in range(octaves):
        lam = 1024 // (4 ** o)
        if lam < 2:
            break
        hr, hc = nrow // lam + 3, ncol // lam + 3
        coarse = torch.rand((1, 1, hr, hc), generator=g).to(dev)
        z = z + (relief * 0.71 * lam) * F.interpolate(coarse, size=(nrow, ncol), mode="bicubic", align_corners=True)[0, 0]
        del coarse
    hb = (rr.to(torch.int64) // 64) * 40503 + (cc.to(torch.int64) // 64) * 9973 + (seed & 0xFFFF)
    hb = (hb * 2654435761) & 0xFFFFFFFF
    hb = ((hb ^ (hb >> 15)) * 2246822519) & 0xFFFFFFFF
    hb = (hb ^ (hb >> 13)) & 0xFFFFFFFF
    hole = hb.to(torch.float32) / 4294967296.0 < nodata_frac
    del hb
    hole[0, :] = True
    hole[-1, :] = True
    hole[:, 0] = True
    hole[:, -1] = True
    return z.to(torch.float64), ~hole


def block_polygons(nrow: int, ncol: int, nby: int, nbx: int, margin: int = 1):
    """Synthetic 'subbasin polygons': an nby x nbx tiling of the grid. Returns (src_cell [nsub],
    poly_off [nsub+1], poly_cells) with 0-based row-major cell ids: src = block centre (the
    point-on-surface cell, Routing:79-81), poly_cells = cells fully inside the block (coverage
    fraction 1, Routing:119-121)."""
    src, off, cells = [], [0], []
    ys = np.linspace(0, nrow, nby + 1).astype(int)
    xs = np.linspace(0, ncol, nbx + 1).astype(int)
    for i in range(nby):
        for j in range(nbx):
            r0, r1, c0, c1 = ys[i] + margin, ys[i + 1] - margin, xs[j] + margin, xs[j + 1] - margin
            src.append(((ys[i] + ys[i + 1]) // 2) * ncol + (xs[j] + xs[j + 1]) // 2)
            if r1 > r0 and c1 > c0:
                rr, cc = np.meshgrid(np.arange(r0, r1), np.arange(c0, c1), indexing="ij")
                cells.append((rr * ncol + cc).ravel())
            else:
                cells.append(np.empty(0, np.int64))
            off.append(off[-1] + len(cells[-1]))
    return (np.asarray(src, np.int32), np.asarray(off, np.int32),
            np.concatenate(cells).astype(np.int32) if cells else np.empty(0, np.int32))
