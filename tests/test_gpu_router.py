"""Month-invariant routing operator (wfa_router_*, the default path of wfa_route) against the dense
sweep and the CPU oracle: same per-polygon maxima, bit for bit."""
import numpy as np
import pytest

from gr2msemidistr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as c
    if c.device_count() < 1:
        pytest.skip("no CUDA device")
    c.set_device(0)
    return c


@pytest.fixture(scope="module")
def oracle():
    import oracle as o
    o.build()
    return o


def same(a, b):
    return np.array_equal(np.nan_to_num(a, nan=-7.0, neginf=-9.0), np.nan_to_num(b, nan=-7.0, neginf=-9.0))


def random_geometry(rng, nrow, ncol, nsub, ncells_per_poly, d8):
    valid = np.flatnonzero((d8.ravel() >= 1) & (d8.ravel() <= 8))
    src = rng.choice(valid, nsub, replace=True).astype(np.int32)
    off = np.zeros(nsub + 1, np.int32)
    cells = []
    for s in range(nsub):
        k = int(rng.integers(0, ncells_per_poly + 1))
        # interior cells: a window around the source plus random cells anywhere (incl. nodata cells)
        r0, c0 = divmod(int(src[s]), ncol)
        rr = np.clip(r0 + rng.integers(-6, 7, k), 0, nrow - 1)
        cc = np.clip(c0 + rng.integers(-6, 7, k), 0, ncol - 1)
        cells.append((rr * ncol + cc).astype(np.int32))
        off[s + 1] = off[s] + k
    return src, off, (np.concatenate(cells) if cells else np.zeros(0, np.int32)).astype(np.int32)


@pytest.mark.parametrize("seed,nrow,ncol,nsub", [(1, 40, 50, 7), (2, 150, 210, 60), (3, 300, 260, 400), (4, 64, 64, 1)])
def test_router_matches_dense_and_oracle(capi, oracle, seed, nrow, ncol, nsub):
    rng = np.random.default_rng(seed)
    d8 = synth.d8_grid(nrow, ncol, seed=seed).numpy()
    src, off, cells = random_geometry(rng, nrow, ncol, nsub, 60, d8)
    QS = rng.uniform(0, 300, (nsub, 37))
    QS[rng.random(QS.shape) < 0.05] = 0.0
    with capi.Plan(d8) as p:
        for fl in (0, capi.WFA_NO_CONTAMINATION, capi.WFA_F32):
            sparse = p.route(QS, src, off, cells, fl)
            dense = p.route(QS, src, off, cells, fl | capi.WFA_ROUTE_DENSE)
            assert same(sparse, dense), fl
        ref = oracle.route(d8, QS, src, off, cells)
        assert same(p.route(QS, src, off, cells), ref)
        with p.router(src, off, cells) as r:
            info = r.info()
            assert 0 < info["nnodes"] <= 2 * nsub and info["nlevels"] >= 1
            # same node forest as the CPU restatement of the operator
            from oracle import router_twin
            op = router_twin.build(d8, src, off, cells)
            assert info["nnodes"] == len(op["nodes"]) and info["nactive"] == op["nactive"]
            assert info["ncandidates"] == sum(len(c) for c in op["cand"])
            assert same(r.apply(QS), ref)
            # columns are just columns: months x parameter sets in one call
            QS2 = rng.uniform(0, 50, (nsub, 3 * 37))
            assert same(r.apply(QS2), p.route(QS2, src, off, cells, capi.WFA_ROUTE_DENSE))


def test_router_edge_cases(capi, oracle):
    rng = np.random.default_rng(9)
    N = -1
    # a 5x6 hand grid: two branches meeting at (2,3), nodata in the corner, everything draining east
    d8 = np.array([[1, 1, 8, N, N, N],
                   [1, 1, 1, 8, 1, 1],
                   [1, 1, 1, 1, 1, 1],
                   [1, 1, 2, 1, 1, 1],
                   [N, 1, 1, 1, 1, 1]], np.int16)
    ncol = 6
    src = np.array([0 * ncol + 0, 3 * ncol + 0, 3 * ncol + 0, 0 * ncol + 4, 2 * ncol + 5], np.int32)  # duplicate + nodata
    off = np.array([0, 3, 3, 8, 9, 11], np.int32)                     # polygon 1 is empty
    cells = np.array([0, 1, 2,  13, 14, 15, 16, 17,  4,  20, 26], np.int32)
    QS = rng.uniform(1, 9, (5, 4))
    with capi.Plan(d8) as p:
        for fl in (0, capi.WFA_NO_CONTAMINATION):
            got = p.route(QS, src, off, cells, fl)
            assert same(got, p.route(QS, src, off, cells, fl | capi.WFA_ROUTE_DENSE))
            assert same(got, oracle.route(d8, QS, src, off, cells, contcheck=not fl))
        assert np.all(np.isneginf(p.route(QS, src, off, cells)[1]))
    # NaN discharge propagates down its path and is skipped by the maximum, as in R (na.rm = TRUE)
    d8 = synth.d8_grid(90, 120, seed=3).numpy()
    src, off, cells = random_geometry(rng, 90, 120, 25, 40, d8)
    QS = rng.uniform(0, 10, (25, 6))
    QS[3, 2] = np.nan
    with capi.Plan(d8) as p:
        assert same(p.route(QS, src, off, cells), p.route(QS, src, off, cells, capi.WFA_ROUTE_DENSE))
        with pytest.raises(capi.GR2MError):
            p.router(np.array([90 * 120], np.int32), np.array([0, 0], np.int32), np.zeros(0, np.int32))


def test_router_large_grid_many_sources(capi):
    # property test at a size the oracle would not finish: operator == dense sweep on a 3000^2 grid
    rng = np.random.default_rng(5)
    n = 3000
    d8 = synth.d8_grid(n, n, seed=2).numpy()
    nsub = 5000
    src, off, cells = random_geometry(rng, n, n, nsub, 200, d8)
    QS = rng.uniform(0, 300, (nsub, 24))
    with capi.Plan(d8) as p:
        with p.router(src, off, cells) as r:
            info = r.info()
            got = r.apply(QS)
        dense = p.route(QS, src, off, cells, capi.WFA_ROUTE_DENSE)
    assert info["nnodes"] >= nsub * 0.9 and info["nactive"] > info["nnodes"]
    assert same(got, dense)


def test_router_degenerate_geometries(capi, oracle):
    rng = np.random.default_rng(11)
    d8 = synth.d8_grid(60, 80, seed=4).numpy()
    nod = np.flatnonzero(~((d8.ravel() >= 1) & (d8.ravel() <= 8)))
    val = np.flatnonzero((d8.ravel() >= 1) & (d8.ravel() <= 8))
    assert nod.size >= 3
    QS = rng.uniform(1, 9, (3, 5))
    with capi.Plan(d8) as p:
        # every source on a nodata cell: nothing is routed, polygons see zeros (or nothing at all)
        src = nod[:3].astype(np.int32)
        off = np.array([0, 4, 4, 6], np.int32)
        cells = np.concatenate([val[:4], nod[:2]]).astype(np.int32)
        with p.router(src, off, cells) as r:
            assert r.info()["nnodes"] == 0 and r.info()["nactive"] == 0
            got = r.apply(QS)
        assert same(got, p.route(QS, src, off, cells, capi.WFA_ROUTE_DENSE))
        assert same(got, oracle.route(d8, QS, src, off, cells))
        # no polygon has an interior cell: all maxima are -Inf, as R's max(numeric(0)) with a warning
        src = val[[5, 50, 500]].astype(np.int32)
        got = p.route(QS, src, np.zeros(4, np.int32), np.zeros(0, np.int32))
        assert np.all(np.isneginf(got))
        # one column, and more columns than one chunk of the dense path would take
        off = np.array([0, 30, 60, 90], np.int32)
        cells = rng.choice(val, 90).astype(np.int32)
        for ncols in (1, 200):
            Q = rng.uniform(0, 50, (3, ncols))
            assert same(p.route(Q, src, off, cells), oracle.route(d8, Q, src, off, cells))
