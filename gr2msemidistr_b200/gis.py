"""Minimal host-side GIS readers and geometry for the Python mirror of the reference's R front end.

The reference leans on raster/terra/sf/geos/exactextractr (R/Routing_GR2MSemiDistr.R:77-81,119-121;
R/Create_Forcing_Inputs.R:95-108); none of them exists in this image, so the once-per-call,
month-invariant geometry the hot path needs is restated here with NumPy only:

  read_tiff            strip TIFF, LZW or uncompressed, float32/int16, chunky or planar bands
  read_shapefile       polygons (.shp) + attribute table (.dbf)
  point_on_surface     GEOS InteriorPointArea (geos_point_on_surface, Routing:80)
  cell_of_points       raster::extract(..., cellnumbers=TRUE) (Routing:81), 0-based row-major
  interior_cells       cells with exactextractr coverage_fraction == 1 (Routing:119-121)
  coverage_fractions   exactextractr coverage_fraction: exact polygon / cell intersection areas
  zonal_weights/means  exact_extract(fun='mean') on the cropped, nearest-neighbour resampled grid
                       (Create:95-108) with those exact fractions

None of this is on the per-month / per-parameter-set path; it runs once per call on the host.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


# ---------------------------------------------------------------------------------------------
# TIFF
# ---------------------------------------------------------------------------------------------
def _lzw_decode(data: bytes, expected: int) -> bytes:
    """TIFF LZW: MSB-first codes, 9..12 bits, early change, Clear = 256, EOI = 257."""
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    bits, nb, pos, width = 0, 0, 0, 9
    prev = None
    n = len(data)
    while True:
        while nb < width and pos < n:
            bits = (bits << 8) | data[pos]
            pos += 1
            nb += 8
        if nb < width:
            break
        code = (bits >> (nb - width)) & ((1 << width) - 1)
        nb -= width
        bits &= (1 << nb) - 1
        if code == 257:
            break
        if code == 256:
            table = table[:258]
            width = 9
            prev = None
            continue
        if prev is None:
            entry = table[code]
        elif code < len(table):
            entry = table[code]
            table.append(prev + entry[:1])
        else:
            entry = prev + prev[:1]
            table.append(entry)
        out += entry
        prev = entry
        if len(table) >= (1 << width) - 1 and width < 12:
            width += 1
        if len(out) >= expected:
            break
    return bytes(out[:expected])


@dataclass
class Raster:
    data: np.ndarray            # [bands, nrow, ncol]
    x0: float                   # west edge
    y0: float                   # north edge
    dx: float
    dy: float                   # positive cell height
    nodata: float | None = None

    @property
    def shape(self):
        return self.data.shape[1:]


def read_tiff(path: str) -> Raster:
    b = open(path, "rb").read()
    bo = "<" if b[:2] == b"II" else ">"
    (off,) = struct.unpack(bo + "I", b[4:8])
    (n,) = struct.unpack(bo + "H", b[off:off + 2])
    tags = {}
    sizes = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 11: 4, 12: 8, 16: 8}
    fmts = {1: "B", 3: "H", 4: "I", 11: "f", 12: "d", 16: "Q"}
    for i in range(n):
        e = b[off + 2 + 12 * i: off + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(bo + "HHI", e[:8])
        sz = sizes.get(typ)
        if sz is None:
            continue
        raw = e[8:8 + sz * cnt] if sz * cnt <= 4 else b[struct.unpack(bo + "I", e[8:12])[0]:][:sz * cnt]
        if typ == 2:
            tags[tag] = raw.split(b"\0")[0].decode(errors="replace")
        elif typ == 5:
            v = struct.unpack(bo + "II" * cnt, raw)
            tags[tag] = tuple(v[2 * j] / v[2 * j + 1] for j in range(cnt))
        else:
            tags[tag] = struct.unpack(bo + fmts[typ] * cnt, raw)
    ncol, nrow = tags[256][0], tags[257][0]
    bps = tags[258][0]
    comp = tags.get(259, (1,))[0]
    spp = tags.get(277, (1,))[0]
    rps = tags.get(278, (nrow,))[0]
    planar = tags.get(284, (1,))[0]
    fmt = tags.get(339, (1,))[0]
    if tags.get(317, (1,))[0] != 1:
        raise ValueError("TIFF predictor not supported")
    dt = {(32, 3): "f4", (64, 3): "f8", (16, 2): "i2", (16, 1): "u2", (32, 2): "i4", (8, 1): "u1"}[(bps, fmt)]
    dtype = np.dtype(bo + dt)
    offs, cnts = tags[273], tags[279]
    strips_per_band = (nrow + rps - 1) // rps
    nplanes = spp if planar == 2 else 1
    samples = 1 if planar == 2 else spp
    out = np.empty((nplanes, nrow, ncol * samples), dtype)
    for p in range(nplanes):
        for s in range(strips_per_band):
            k = p * strips_per_band + s
            r0 = s * rps
            rows = min(rps, nrow - r0)
            need = rows * ncol * samples * dtype.itemsize
            raw = b[offs[k]:offs[k] + cnts[k]]
            if comp == 5:
                raw = _lzw_decode(raw, need)
            elif comp != 1:
                raise ValueError(f"TIFF compression {comp} not supported")
            out[p, r0:r0 + rows] = np.frombuffer(raw[:need], dtype).reshape(rows, ncol * samples)
    data = out if planar == 2 else out.reshape(nrow, ncol, spp).transpose(2, 0, 1)
    scale, tie = tags[33550], tags[33922]
    nod = tags.get(42113)
    return Raster(np.ascontiguousarray(data).astype(dtype.newbyteorder("=")), tie[3] - tie[0] * scale[0],
                  tie[4] + tie[1] * scale[1], scale[0], scale[1], float(nod) if nod else None)


# ---------------------------------------------------------------------------------------------
# shapefile (+ dbf)
# ---------------------------------------------------------------------------------------------
@dataclass
class Subbasins:
    """What the reference reads from the SpatialPolygonsDataFrame: @data$Area, @data$Region, $COMID
    (R/Run_GR2MSemiDistr.R:68-71) and the polygons (R/Routing_GR2MSemiDistr.R:79-80)."""
    Area: np.ndarray
    Region: np.ndarray
    COMID: np.ndarray
    polygons: list = field(default_factory=list)     # per subbasin: list of rings, each [n, 2] (x, y)

    def __len__(self):
        return len(self.Area)


def read_dbf(path: str):
    b = open(path, "rb").read()
    nrec, hlen, rlen = struct.unpack("<IHH", b[4:12])
    fields, p = [], 32
    while b[p] != 0x0D:
        name = b[p:p + 11].split(b"\0")[0].decode()
        fields.append((name, chr(b[p + 11]), b[p + 16]))
        p += 32
    cols = {f[0]: [] for f in fields}
    for r in range(nrec):
        rec = b[hlen + r * rlen: hlen + (r + 1) * rlen]
        q = 1
        for name, typ, ln in fields:
            v = rec[q:q + ln].decode("latin-1").strip()
            q += ln
            cols[name].append(float(v) if typ in "NF" and v else v)
    return cols


def read_shp(path: str):
    b = open(path, "rb").read()
    pos, polys = 100, []
    while pos < len(b):
        (clen,) = struct.unpack(">i", b[pos + 4:pos + 8])
        rec = b[pos + 8: pos + 8 + 2 * clen]
        pos += 8 + 2 * clen
        (stype,) = struct.unpack("<i", rec[:4])
        if stype not in (5, 15, 25):
            polys.append([])
            continue
        nparts, npts = struct.unpack("<ii", rec[36:44])
        parts = list(struct.unpack(f"<{nparts}i", rec[44:44 + 4 * nparts])) + [npts]
        pts = np.frombuffer(rec[44 + 4 * nparts: 44 + 4 * nparts + 16 * npts], "<f8").reshape(npts, 2)
        polys.append([pts[parts[i]:parts[i + 1]].copy() for i in range(nparts)])
    return polys


def read_shapefile(stem: str) -> Subbasins:
    a = read_dbf(stem + ".dbf")
    return Subbasins(np.asarray(a["Area"], float), np.asarray(a["Region"]), np.asarray(a["COMID"]).astype(int),
                     read_shp(stem + ".shp"))


# ---------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------
def points_in_polygon(x, y, rings) -> np.ndarray:
    """Even-odd rule over all rings (holes included), vectorised over the query points."""
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    inside = np.zeros(x.shape, bool)
    for ring in rings:
        x1, y1 = ring[:-1, 0], ring[:-1, 1]
        x2, y2 = ring[1:, 0], ring[1:, 1]
        for a, b, c, d in zip(x1, y1, x2, y2):
            if b == d:
                continue
            cond = (b > y) != (d > y)
            xi = a + (y - b) * (c - a) / (d - b)
            inside ^= cond & (x < xi)
    return inside


def ring_area(ring) -> float:
    """Signed shoelace area: positive for counter-clockwise rings (shapefile holes), negative for clockwise
    ones (shapefile outer rings)."""
    x, y = ring[:, 0], ring[:, 1]
    return 0.5 * float(np.sum(x[:-1] * y[1:] - x[1:] * y[:-1]))


def split_parts(rings):
    """Group the rings of a (multi-part) shapefile polygon into [outer, hole, ...] lists: clockwise rings are
    shells, every other ring is a hole of the shell that contains its first vertex."""
    rings = [r for r in rings if len(r) >= 4]
    shells = [r for r in rings if ring_area(r) < 0]
    if not shells:                       # orientation not as in the shapefile spec: one part, even-odd semantics
        return [list(rings)] if rings else []
    parts = [[r] for r in shells]
    for r in rings:
        if ring_area(r) >= 0:
            for part in parts:
                if points_in_polygon(np.array([r[0, 0]]), np.array([r[0, 1]]), [part[0]])[0]:
                    part.append(r)
                    break
    return parts


def _interior_point_of_part(rings):
    """GEOS InteriorPointArea for one polygon (shell + holes): the scan line runs midway between the two
    vertex ordinates that bracket the centre of the envelope, the point is the middle of the widest crossing
    interval.  Returns (x, y, width)."""
    ys = np.concatenate([r[:, 1] for r in rings])
    cy = 0.5 * (ys.min() + ys.max())
    lo = ys[ys <= cy].max() if np.any(ys <= cy) else ys.min()
    hi = ys[ys > cy].min() if np.any(ys > cy) else ys.max()
    sy = 0.5 * (lo + hi)
    xs = []
    for r in rings:
        for (a, b), (c, d) in zip(r[:-1], r[1:]):
            if (b > sy) != (d > sy):
                xs.append(a + (sy - b) * (c - a) / (d - b))
    xs.sort()
    best, bx = -1.0, float(np.mean([r[:, 0].mean() for r in rings]))
    for i in range(0, len(xs) - 1, 2):
        w = xs[i + 1] - xs[i]
        if w > best:
            best, bx = w, 0.5 * (xs[i] + xs[i + 1])
    return bx, sy, best


def point_on_surface(rings):
    """geos_point_on_surface (Routing:80) = GEOS InteriorPointArea: every part of a multi-part polygon gets its
    own scan line, the part with the widest interior section wins."""
    best = None
    for part in split_parts(rings):
        x, y, w = _interior_point_of_part(part)
        if best is None or w > best[2]:
            best = (x, y, w)
    if best is None:
        raise ValueError("polygon without rings")
    return best[0], best[1]


def cell_of_points(r: Raster, xs, ys) -> np.ndarray:
    """0-based row-major cell containing each point, -1 outside (raster::extract cellnumbers - 1)."""
    nrow, ncol = r.shape
    col = np.floor((np.asarray(xs) - r.x0) / r.dx).astype(int)
    row = np.floor((r.y0 - np.asarray(ys)) / r.dy).astype(int)
    ok = (col >= 0) & (col < ncol) & (row >= 0) & (row < nrow)
    return np.where(ok, row * ncol + col, -1).astype(np.int32)


def _clip_halfplane(pts, axis, bound, keep_greater):
    """Sutherland-Hodgman against one axis-aligned half plane.  pts: open ring [n, 2] (no repeated last vertex)."""
    if len(pts) == 0:
        return pts
    v = pts[:, axis]
    inside = (v >= bound) if keep_greater else (v <= bound)
    nxt = np.roll(pts, -1, axis=0)
    nin = np.roll(inside, -1)
    cross = inside != nin
    out = []
    if cross.any():
        d = nxt[:, axis] - v
        t = np.where(cross, (bound - v) / np.where(d == 0.0, 1.0, d), 0.0)
        ip = pts + t[:, None] * (nxt - pts)
        ip[:, axis] = bound
    for i in range(len(pts)):
        if inside[i]:
            out.append(pts[i])
        if cross[i]:
            out.append(ip[i])
    return np.array(out).reshape(-1, 2)


def _open_area(pts) -> float:
    if len(pts) < 3:
        return 0.0
    x, y = pts[:, 0], pts[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _boundary_cells(rings, x0, y0, dx, dy, ncol, nrow):
    """Boolean [nrow, ncol]: cells whose interior a polygon edge passes through -- exact grid traversal (every
    crossing of a grid line splits the edge; the midpoint of each piece names its cell)."""
    hit = np.zeros((nrow, ncol), bool)
    for g in rings:
        for (a, b), (c, d) in zip(g[:-1], g[1:]):
            ts = [0.0, 1.0]
            if c != a:
                k0, k1 = sorted(((a - x0) / dx, (c - x0) / dx))
                ks = np.arange(np.ceil(k0), np.floor(k1) + 1)
                ts.extend(((x0 + ks * dx - a) / (c - a)).tolist())
            if d != b:
                k0, k1 = sorted(((y0 - b) / dy, (y0 - d) / dy))
                ks = np.arange(np.ceil(k0), np.floor(k1) + 1)
                ts.extend(((y0 - ks * dy - b) / (d - b)).tolist())
            t = np.unique(np.clip(np.array(ts), 0.0, 1.0))
            tm = 0.5 * (t[:-1] + t[1:])
            cc = np.floor((a + tm * (c - a) - x0) / dx).astype(int)
            rr = np.floor((y0 - (b + tm * (d - b))) / dy).astype(int)
            ok = (cc >= 0) & (cc < ncol) & (rr >= 0) & (rr < nrow)
            hit[rr[ok], cc[ok]] = True
    return hit


def coverage_fractions(rings, x0, y0, dx, dy, ncol, nrow):
    """exactextractr's coverage_fraction on the grid with north-west corner (x0, y0): the exact planar share of
    every cell's area that lies inside the polygon (holes and multiple parts by signed ring area).  Cells no
    edge passes through are 1 or 0 by their centre; the others are clipped (Sutherland-Hodgman, first to the
    row, then to the cell).  Returns float64 [nrow, ncol]."""
    cov = np.zeros((nrow, ncol))
    if not rings:
        return cov
    xs = np.concatenate([g[:, 0] for g in rings])
    ys = np.concatenate([g[:, 1] for g in rings])
    c0 = max(0, int(np.floor((xs.min() - x0) / dx)))
    c1 = min(ncol, int(np.floor((xs.max() - x0) / dx)) + 1)
    r0 = max(0, int(np.floor((y0 - ys.max()) / dy)))
    r1 = min(nrow, int(np.floor((y0 - ys.min()) / dy)) + 1)
    if c1 <= c0 or r1 <= r0:
        return cov
    cx = x0 + dx * (np.arange(c0, c1) + 0.5)
    cy = y0 - dy * (np.arange(r0, r1) + 0.5)
    X, Y = np.meshgrid(cx, cy)
    cov[r0:r1, c0:c1] = points_in_polygon(X, Y, rings)
    bnd = _boundary_cells(rings, x0, y0, dx, dy, ncol, nrow)
    open_rings = [np.asarray(g[:-1], float) if np.array_equal(g[0], g[-1]) else np.asarray(g, float) for g in rings]
    cell_area = dx * dy
    for rr in np.flatnonzero(bnd.any(axis=1)):
        ytop, ybot = y0 - rr * dy, y0 - (rr + 1) * dy
        strips = [_clip_halfplane(_clip_halfplane(g, 1, ybot, True), 1, ytop, False) for g in open_rings]
        for cc in np.flatnonzero(bnd[rr]):
            xl, xr = x0 + cc * dx, x0 + (cc + 1) * dx
            a = sum(_open_area(_clip_halfplane(_clip_halfplane(g, 0, xl, True), 0, xr, False)) for g in strips)
            cov[rr, cc] = min(1.0, max(0.0, abs(a) / cell_area))
    return cov


def interior_cells(r: Raster, rings) -> np.ndarray:
    """Cells with exactextractr coverage_fraction == 1 (Routing:119-121): centre inside the polygon and no
    polygon edge through the cell's interior."""
    nrow, ncol = r.shape
    if not rings:
        return np.empty(0, np.int32)
    xs = np.concatenate([g[:, 0] for g in rings])
    ys = np.concatenate([g[:, 1] for g in rings])
    c0 = max(0, int(np.floor((xs.min() - r.x0) / r.dx)))
    c1 = min(ncol, int(np.floor((xs.max() - r.x0) / r.dx)) + 1)
    r0 = max(0, int(np.floor((r.y0 - ys.max()) / r.dy)))
    r1 = min(nrow, int(np.floor((r.y0 - ys.min()) / r.dy)) + 1)
    if c1 <= c0 or r1 <= r0:
        return np.empty(0, np.int32)
    X, Y = np.meshgrid(r.x0 + r.dx * (np.arange(c0, c1) + 0.5), r.y0 - r.dy * (np.arange(r0, r1) + 0.5))
    inside = points_in_polygon(X, Y, rings)
    bnd = _boundary_cells(rings, r.x0, r.y0, r.dx, r.dy, ncol, nrow)[r0:r1, c0:c1]
    rr, cc = np.nonzero(inside & ~bnd)
    return ((rr + r0) * ncol + (cc + c0)).astype(np.int32)


def crop_window(r: Raster, rings_list, buffer: float = 1.1):
    """raster::crop(x, extent(roi) * Buffer) (Create:95): the polygons' bounding box scaled about its centre,
    snapped to the nearest cell edges, intersected with the raster.  Returns (row0, row1, col0, col1)."""
    xs = np.concatenate([g[:, 0] for rings in rings_list for g in rings])
    ys = np.concatenate([g[:, 1] for rings in rings_list for g in rings])
    nrow, ncol = r.shape
    hx, hy = 0.5 * buffer * (xs.max() - xs.min()), 0.5 * buffer * (ys.max() - ys.min())
    mx, my = 0.5 * (xs.max() + xs.min()), 0.5 * (ys.max() + ys.min())
    c0 = int(np.floor((mx - hx - r.x0) / r.dx + 0.5))
    c1 = int(np.floor((mx + hx - r.x0) / r.dx + 0.5))
    r0 = int(np.floor((r.y0 - (my + hy)) / r.dy + 0.5))
    r1 = int(np.floor((r.y0 - (my - hy)) / r.dy + 0.5))
    return max(r0, 0), min(r1, nrow), max(c0, 0), min(c1, ncol)


def zonal_weights(r: Raster, rings, resolution: float = 0.01, window=None) -> np.ndarray:
    """Weight [nrow, ncol] of every source cell in exact_extract(resample(x, fine, 'ngb'), polygon, 'mean')
    (Create:96-106): the summed coverage fractions of the `resolution` cells that take its value, i.e. the
    area of polygon n fine cell in units of the fine cell.  `window` = (row0, row1, col0, col1) of the cropped
    brick (crop_window); the fine grid starts at its north-west corner.  When the fine grid nests in the source
    grid (0.1 / 0.01) the sum over a source cell is the exact area of polygon n source cell, so the source grid
    is clipped directly; otherwise the fine grid is clipped and each fine cell credited to its nearest source cell."""
    nrow, ncol = r.shape
    r0, r1, c0, c1 = window if window is not None else (0, nrow, 0, ncol)
    w = np.zeros((nrow, ncol))
    if r1 <= r0 or c1 <= c0:
        return w
    wx0, wy0 = r.x0 + c0 * r.dx, r.y0 - r0 * r.dy
    kx, ky = r.dx / resolution, r.dy / resolution
    if abs(kx - round(kx)) < 1e-6 and abs(ky - round(ky)) < 1e-6 and round(kx) >= 1 and round(ky) >= 1:
        cov = coverage_fractions(rings, wx0, wy0, r.dx, r.dy, c1 - c0, r1 - r0)
        w[r0:r1, c0:c1] = cov * (round(kx) * round(ky))
        return w
    nfx, nfy = int(np.ceil((c1 - c0) * r.dx / resolution - 1e-9)), int(np.ceil((r1 - r0) * r.dy / resolution - 1e-9))
    cov = coverage_fractions(rings, wx0, wy0, resolution, resolution, nfx, nfy)
    fr, fc = np.nonzero(cov)
    cells = cell_of_points(r, wx0 + (fc + 0.5) * resolution, wy0 - (fr + 0.5) * resolution)
    ok = cells >= 0
    np.add.at(w.reshape(-1), cells[ok], cov[fr[ok], fc[ok]])
    return w


def zonal_means(r: Raster, subb: Subbasins, resolution: float = 0.01, buffer: float | None = 1.1) -> np.ndarray:
    """[nsub, nbands] exact_extract(fun='mean'): coverage-weighted means ignoring nodata cells; rounded like
    Create:114 by the caller.  buffer=None skips the crop."""
    out = np.empty((len(subb), r.data.shape[0]))
    valid = np.ones(r.data.shape, bool) if r.nodata is None else ~np.isclose(r.data, r.nodata, rtol=1e-6)
    valid &= np.isfinite(r.data)
    window = crop_window(r, subb.polygons, buffer) if buffer is not None else None
    for i, rings in enumerate(subb.polygons):
        w = zonal_weights(r, rings, resolution, window)[None]
        ww = w * valid
        den = ww.sum((1, 2))
        out[i] = np.where(den > 0, (ww * np.where(valid, r.data, 0.0)).sum((1, 2)) / np.maximum(den, 1e-300), np.nan)
    return out
