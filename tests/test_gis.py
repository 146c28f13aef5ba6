"""Host-side geometry used once per call by the routing / forcing wrappers (gis.py)."""
import numpy as np

from gr2msemidistr_b200 import gis


def _raster(nrow=10, ncol=12, dx=1.0):
    return gis.Raster(np.zeros((1, nrow, ncol), np.float32), x0=100.0, y0=50.0, dx=dx, dy=dx)


def test_points_in_polygon_with_hole():
    outer = np.array([[0, 0], [10, 0], [10, 10], [0, 10], [0, 0]], float)
    hole = np.array([[4, 4], [6, 4], [6, 6], [4, 6], [4, 4]], float)
    x = np.array([1.0, 5.0, 11.0, 9.9, 4.5])
    y = np.array([1.0, 5.0, 5.0, 9.9, 3.5])
    assert gis.points_in_polygon(x, y, [outer, hole]).tolist() == [True, False, False, True, True]


def test_point_on_surface_is_inside_even_for_concave_shapes():
    # a U shape: the centroid falls in the notch, the point on surface must not
    u = np.array([[0, 0], [9, 0], [9, 9], [6, 9], [6, 3], [3, 3], [3, 9], [0, 9], [0, 0]], float)
    px, py = gis.point_on_surface([u])
    assert gis.points_in_polygon(np.array([px]), np.array([py]), [u])[0]
    cx, cy = u[:-1].mean(0)
    assert not gis.points_in_polygon(np.array([4.5]), np.array([6.0]), [u])[0]      # the notch is outside
    sq = np.array([[2, 2], [4, 2], [4, 6], [2, 6], [2, 2]], float)
    assert gis.point_on_surface([sq]) == (3.0, 4.0)


def test_cell_of_points_row_major_zero_based():
    r = _raster()
    # north-west corner cell is 0; one row down adds ncol (raster cell numbers are row-major from the top)
    cells = gis.cell_of_points(r, [100.5, 111.5, 100.5, 99.0], [49.5, 49.5, 48.5, 45.0])
    assert cells.tolist() == [0, 11, 12, -1]


def test_interior_cells_are_the_fully_covered_ones():
    r = _raster()
    # polygon covering x in [102.5, 107.5], y in [43.5, 47.5]: full cells are cols 3..6, rows 3..5
    p = np.array([[102.5, 43.5], [107.5, 43.5], [107.5, 47.5], [102.5, 47.5], [102.5, 43.5]])
    got = sorted(gis.interior_cells(r, [p]).tolist())
    want = sorted(rr * 12 + cc for rr in (3, 4, 5) for cc in (3, 4, 5, 6))
    assert got == want
    tiny = np.array([[103.2, 46.2], [103.8, 46.2], [103.8, 46.8], [103.2, 46.8], [103.2, 46.2]])
    assert gis.interior_cells(r, [tiny]).size == 0                                   # max() of nothing -> -Inf in R


def test_zonal_mean_weights_follow_coverage():
    r = gis.Raster(np.stack([np.arange(12.0).reshape(3, 4), np.ones((3, 4))]).astype(np.float32), 0.0, 3.0, 1.0, 1.0,
                   nodata=-9999.0)
    r.data[0, 0, 0] = -9999.0
    poly = np.array([[0, 1], [2, 1], [2, 3], [0, 3], [0, 1]], float)      # cells (0,0) (0,1) (1,0) (1,1)
    sb = gis.Subbasins(np.array([1.0]), np.array(["A"]), np.array([1]), [[poly]])
    m = gis.zonal_means(r, sb, resolution=0.1)
    assert abs(m[0, 0] - (1 + 4 + 5) / 3) < 1e-9 and abs(m[0, 1] - 1.0) < 1e-12    # nodata cell ignored
    w = gis.zonal_weights(r, [poly], 0.1)
    assert w[0, 0] == w[1, 1] == 100 and w[2, 3] == 0


def test_coverage_fractions_hand_computed():
    # unit grid, north-west corner (0, 4), 4 x 4 cells; row 0 is y in [3, 4]
    def cov(*rings):
        return gis.coverage_fractions([np.array(r, float) for r in rings], 0.0, 4.0, 1.0, 1.0, 4, 4)
    # axis-aligned rectangle [0.5, 2.25] x [1, 3.5]
    c = cov([[0.5, 1], [2.25, 1], [2.25, 3.5], [0.5, 3.5], [0.5, 1]])
    want = np.zeros((4, 4))
    want[0, :3] = [0.25, 0.5, 0.125]          # y in [3, 3.5]: half a row
    want[1, :3] = want[2, :3] = [0.5, 1.0, 0.25]
    assert np.allclose(c, want, atol=1e-15)
    # right triangle (0,0) (4,0) (0,4): cells on the diagonal are half covered, below it full
    t = cov([[0, 0], [4, 0], [0, 4], [0, 0]])
    for r in range(4):
        for col in range(4):
            below = col + (3 - r) < 3
            on = col + (3 - r) == 3
            assert abs(t[r, col] - (1.0 if below else 0.5 if on else 0.0)) < 1e-15
    assert abs(t.sum() - 8.0) < 1e-14
    # clockwise orientation gives the same fractions
    assert np.allclose(cov([[0, 0], [0, 4], [4, 0], [0, 0]]), t, atol=1e-15)
    # a hole (opposite orientation) is subtracted; a hole strictly inside one cell leaves that cell partial
    h = cov([[0, 0], [0, 4], [4, 4], [4, 0], [0, 0]], [[1.25, 1.25], [1.75, 1.25], [1.75, 1.75], [1.25, 1.75], [1.25, 1.25]])
    assert abs(h[2, 1] - 0.75) < 1e-15 and abs(h.sum() - 15.75) < 1e-14
    # two parts
    m = cov([[0, 0], [0, 1.5], [1.5, 1.5], [1.5, 0], [0, 0]], [[3, 3], [3, 4], [4, 4], [4, 3], [3, 3]])
    assert abs(m.sum() - 3.25) < 1e-14 and m[0, 3] == 1.0 and abs(m[2, 1] - 0.25) < 1e-15
    # polygon reaching outside the grid: only the part on the grid counts
    o = cov([[-2, -2], [2.5, -2], [2.5, 0.5], [-2, 0.5], [-2, -2]])
    assert np.allclose(o[3], [0.5, 0.5, 0.25, 0.0]) and o[:3].sum() == 0.0


def test_coverage_fractions_match_dense_sampling(rng=np.random.default_rng(3)):
    # an irregular concave polygon against 400 x 400 sub-samples per cell
    ang = np.sort(rng.uniform(0, 2 * np.pi, 23))
    rad = rng.uniform(1.0, 3.4, 23)
    ring = np.column_stack([3.6 + rad * np.cos(ang), 3.4 + rad * np.sin(ang)])
    ring = np.vstack([ring, ring[:1]])
    c = gis.coverage_fractions([ring], 0.0, 7.0, 1.0, 1.0, 8, 7)
    assert abs(c.sum() - abs(gis.ring_area(ring))) < 1e-12            # fractions add up to the polygon's area
    n = 400
    u = (np.arange(n) + 0.5) / n
    for (r, col) in [(3, 3), (1, 4), (5, 2), (3, 6), (0, 3)]:
        X, Y = np.meshgrid(col + u, 7.0 - r - u)
        assert abs(gis.points_in_polygon(X, Y, [ring]).mean() - c[r, col]) < 5e-3
    inner = gis.interior_cells(gis.Raster(np.zeros((1, 7, 8), np.float32), 0.0, 7.0, 1.0, 1.0), [ring])
    assert sorted(inner.tolist()) == sorted(np.flatnonzero(c.reshape(-1) == 1.0).tolist())


def test_zonal_weights_nested_and_general_grids_agree():
    r = gis.Raster(np.zeros((1, 6, 6), np.float32), 10.0, 6.0, 1.0, 1.0)
    ring = np.array([[10.3, 0.7], [15.2, 1.1], [14.1, 5.6], [11.9, 4.2], [10.3, 0.7]])
    nested = gis.zonal_weights(r, [ring], 0.1)                         # 10 x 10 fine cells per source cell
    assert abs(nested.sum() / 100.0 - abs(gis.ring_area(ring))) < 1e-12
    # same fine grid through the general (non-nested) branch: 0.1 * (1 + 1e-5) does not divide the cell size
    general = gis.zonal_weights(r, [ring], 0.1 * (1 + 1e-5))
    assert np.max(np.abs(general - nested)) < 0.25                     # < a quarter fine cell anywhere
    # crop window (Create:95): cells outside the window carry no weight
    win = gis.crop_window(r, [[ring]], 1.0)                            # bbox snapped to the nearest cell edges
    assert win == (0, 5, 0, 5)
    cropped = gis.zonal_weights(r, [ring], 0.1, win)
    assert cropped[:, 5].sum() == 0.0 and cropped[5].sum() == 0.0 and np.array_equal(cropped[:5, :5], nested[:5, :5])


def test_point_on_surface_multipart_uses_the_widest_part():
    # shapefile orientation: shells clockwise, holes counter-clockwise
    small = np.array([[0, 0], [0, 1], [1, 1], [1, 0], [0, 0]], float)
    big = np.array([[10, 0], [10, 6], [18, 6], [18, 0], [10, 0]], float)
    hole = np.array([[12, 2], [16, 2], [16, 4], [12, 4], [12, 2]], float)
    assert gis.ring_area(small) < 0 < gis.ring_area(hole)
    parts = gis.split_parts([small, big, hole])
    assert len(parts) == 2 and len(parts[0]) == 1 and len(parts[1]) == 2
    x, y = gis.point_on_surface([small, big, hole])
    # the big part's scan line (y = 3: midway between the vertex ordinates 2 and 4 that bracket the envelope's
    # centre) crosses 10..12 and 16..18, both 2 wide -- wider than anything in the small part
    assert gis.points_in_polygon(np.array([x]), np.array([y]), [big, hole])[0]
    assert (x, y) == (11.0, 3.0)
