"""Minimal host-side GIS readers and geometry for the Python mirror of the reference's R front end.

The reference leans on raster/terra/sf/geos/exactextractr (R/Routing_GR2MSemiDistr.R:77-81,119-121;
R/Create_Forcing_Inputs.R:95-108); none of them exists in this image, so the once-per-call,
month-invariant geometry the hot path needs is restated here with NumPy only:

  read_tiff            strip TIFF, LZW or uncompressed, float32/int16, chunky or planar bands
  read_shapefile       polygons (.shp) + attribute table (.dbf)
  point_on_surface     GEOS InteriorPointArea (geos_point_on_surface, Routing:80)
  cell_of_points       raster::extract(..., cellnumbers=TRUE) (Routing:81), 0-based row-major
  interior_cells       cells with exactextractr coverage_fraction == 1 (Routing:119-121)
  zonal_weights/means  exact_extract(fun='mean') on the nearest-neighbour resampled grid
                       (Create:95-108), approximated by cell-centre sampling on that grid

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


def point_on_surface(rings):
    """GEOS InteriorPointArea: scan line through the middle of the widest interior strip at a y that
    avoids vertices, midpoint of the widest crossing interval."""
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
    return bx, sy


def cell_of_points(r: Raster, xs, ys) -> np.ndarray:
    """0-based row-major cell containing each point, -1 outside (raster::extract cellnumbers - 1)."""
    nrow, ncol = r.shape
    col = np.floor((np.asarray(xs) - r.x0) / r.dx).astype(int)
    row = np.floor((r.y0 - np.asarray(ys)) / r.dy).astype(int)
    ok = (col >= 0) & (col < ncol) & (row >= 0) & (row < nrow)
    return np.where(ok, row * ncol + col, -1).astype(np.int32)


def interior_cells(r: Raster, rings) -> np.ndarray:
    """Cells entirely inside the polygon (coverage fraction 1): all four corners inside and no
    polygon edge passing through the cell."""
    nrow, ncol = r.shape
    xs = np.concatenate([g[:, 0] for g in rings])
    ys = np.concatenate([g[:, 1] for g in rings])
    c0 = max(0, int(np.floor((xs.min() - r.x0) / r.dx)))
    c1 = min(ncol, int(np.ceil((xs.max() - r.x0) / r.dx)))
    r0 = max(0, int(np.floor((r.y0 - ys.max()) / r.dy)))
    r1 = min(nrow, int(np.ceil((r.y0 - ys.min()) / r.dy)))
    if c1 <= c0 or r1 <= r0:
        return np.empty(0, np.int32)
    gx = r.x0 + r.dx * np.arange(c0, c1 + 1)
    gy = r.y0 - r.dy * np.arange(r0, r1 + 1)
    X, Y = np.meshgrid(gx, gy)
    corner = points_in_polygon(X, Y, rings)
    full = corner[:-1, :-1] & corner[1:, :-1] & corner[:-1, 1:] & corner[1:, 1:]
    touched = np.zeros_like(full)
    step = 0.2 * min(r.dx, r.dy)
    for g in rings:
        for (a, b), (c, d) in zip(g[:-1], g[1:]):
            n = max(2, int(np.hypot(c - a, d - b) / step) + 2)
            t = np.linspace(0.0, 1.0, n)
            cc = np.floor((a + t * (c - a) - r.x0) / r.dx).astype(int) - c0
            rr = np.floor((r.y0 - (b + t * (d - b))) / r.dy).astype(int) - r0
            ok = (cc >= 0) & (cc < full.shape[1]) & (rr >= 0) & (rr < full.shape[0])
            touched[rr[ok], cc[ok]] = True
    rr, cc = np.nonzero(full & ~touched)
    return ((rr + r0) * ncol + (cc + c0)).astype(np.int32)


def zonal_weights(r: Raster, rings, resolution: float = 0.01) -> np.ndarray:
    """Weights [nrow, ncol] of each source cell in the polygon mean, from sampling the polygon on the
    `resolution` grid the reference resamples to with method 'ngb' (Create:96-98,105)."""
    xs = np.concatenate([g[:, 0] for g in rings])
    ys = np.concatenate([g[:, 1] for g in rings])
    fx = np.arange(np.floor(xs.min() / resolution), np.ceil(xs.max() / resolution) + 1) * resolution + resolution / 2
    fy = np.arange(np.floor(ys.min() / resolution), np.ceil(ys.max() / resolution) + 1) * resolution + resolution / 2
    X, Y = np.meshgrid(fx, fy)
    inside = points_in_polygon(X, Y, rings)
    cells = cell_of_points(r, X[inside], Y[inside])
    w = np.bincount(cells[cells >= 0], minlength=r.shape[0] * r.shape[1]).astype(float)
    return w.reshape(r.shape)


def zonal_means(r: Raster, subb: Subbasins, resolution: float = 0.01) -> np.ndarray:
    """[nsub, nbands] coverage-weighted means ignoring nodata, rounded like Create:114 by the caller."""
    out = np.empty((len(subb), r.data.shape[0]))
    valid = np.ones(r.data.shape, bool) if r.nodata is None else ~np.isclose(r.data, r.nodata, rtol=1e-6)
    valid &= np.isfinite(r.data)
    for i, rings in enumerate(subb.polygons):
        w = zonal_weights(r, rings, resolution)[None]
        ww = w * valid
        out[i] = (ww * np.where(valid, r.data, 0.0)).sum((1, 2)) / np.maximum(ww.sum((1, 2)), 1e-300)
    return out
