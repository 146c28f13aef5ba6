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
