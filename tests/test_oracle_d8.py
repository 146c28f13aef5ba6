"""TauDEM AreaD8 -wg restatement and the level plan (Routing:100-123)."""
import numpy as np

from oracle import twin

N = -32768


def test_3x3_confluence_hand(oracle):
    # all cells valid, everything drains to the centre, centre drains east (off-grid at (1,2)->(1,3))
    # codes: 1=E 2=NE 3=N 4=NW 5=W 6=SW 7=S 8=SE
    d = np.array([[8, 7, 6],
                  [1, 1, 5],
                  [2, 3, 4]], np.int16)
    W = np.arange(1.0, 10.0).reshape(3, 3)
    A = oracle.aread8(d, W, contcheck=False)
    # (1,2) drains W into the centre (code 5) and the centre drains E into (1,2): a 2-cycle. Kahn
    # never evaluates cycle cells (they stay nodata, as in TauDEM); the 7 heads evaluate normally.
    assert np.isnan(A[1, 1]) and np.isnan(A[1, 2])
    heads = [(0, 0), (0, 1), (0, 2), (1, 0), (2, 0), (2, 1), (2, 2)]
    assert all(A[r, c] == W[r, c] for r, c in heads)


def test_5x5_hand_with_nodata_and_contamination(oracle):
    d = np.array([[N, N, N, N, N],
                  [N, 8, 7, 6, N],
                  [N, 1, 7, 5, N],
                  [N, 1, 7, 5, N],
                  [N, N, N, N, N]], np.int16)
    W = np.zeros((5, 5)); W[1:4, 1:4] = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    A = oracle.aread8(d, W, contcheck=False)
    # (2,2) gets k=1 E (2,3), k=2 NE (1,3), k=3 N (1,2), k=4 NW (1,1), k=5 W (2,1) in that order
    assert A[2, 2] == 5 + 6 + 3 + 2 + 1 + 4
    assert A[3, 2] == 8 + 9 + A[2, 2] + 7        # k=1 E (3,3), k=3 N (2,2), k=5 W (3,1)
    assert np.isnan(A[0, 0]) and A[1, 1] == 1
    # with contamination every cell touches nodata except (2,2); (2,2) receives contaminated cells
    Ac = oracle.aread8(d, W, contcheck=True)
    assert np.all(np.isnan(Ac))
    pl = oracle.d8_plan(d)
    assert pl["nlevels"] == 3
    assert pl["level"].reshape(5, 5)[2, 2] == 1 and pl["level"].reshape(5, 5)[3, 2] == 2
    assert pl["receiver"].reshape(5, 5)[3, 2] == -1          # drains into nodata
    assert pl["indeg"].reshape(5, 5)[2, 2] == 5 and pl["inmask"].reshape(5, 5)[2, 2] == 0b00011111
    assert list(pl["order"][-2:]) == [2 * 5 + 2, 3 * 5 + 2]
    assert pl["contam"].reshape(5, 5)[1:4, 1:4].all()


def test_contamination_only_downstream_of_edges(oracle):
    from gr2msemidistr_b200 import synth
    d = synth.d8_grid(120, 160).numpy()
    pl = oracle.d8_plan(d)
    A = oracle.aread8(d, np.ones(d.shape), contcheck=True)
    valid = (d >= 1) & (d <= 8)
    assert np.array_equal(np.isnan(A).ravel(), ~valid.ravel() | (pl["contam"] == 1))
    A0 = oracle.aread8(d, np.ones(d.shape), contcheck=False)
    ok = ~np.isnan(A)
    assert np.array_equal(A[ok], A0[ok]) and 0 < ok.sum() < valid.sum()


def test_plan_invariants_and_twin(oracle, rng):
    from gr2msemidistr_b200 import synth
    d = synth.d8_grid(90, 130).numpy()
    W = rng.uniform(0, 100, d.shape)
    pl = oracle.d8_plan(d)
    valid = ((d >= 1) & (d <= 8)).ravel()
    lev, rec, order, off = pl["level"], pl["receiver"], pl["order"], pl["level_off"]
    assert len(order) == valid.sum() and np.array_equal(np.sort(order), np.flatnonzero(valid))
    assert np.all(lev[~valid] == -1) and np.all(lev[valid] >= 0)
    has = rec >= 0
    assert np.all(lev[rec[has]] > lev[has])
    assert np.array_equal(pl["indeg"], np.bincount(rec[has], minlength=d.size))
    # order is sorted by (level, cell id); level_off delimits the levels
    key = lev[order].astype(np.int64) * d.size + order
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(np.bincount(lev[valid]), np.diff(off))
    for f32 in (False, True):
        A = oracle.aread8(d, W, f32=f32, contcheck=False)
        At, Lt = twin.aread8(d, W, contcheck=False, f32=f32)
        assert np.array_equal(np.nan_to_num(A, nan=-1), np.nan_to_num(At, nan=-1))
        assert np.array_equal(Lt.ravel(), lev)
    Ac = oracle.aread8(d, W, contcheck=True)
    Atc, _ = twin.aread8(d, W, contcheck=True)
    assert np.array_equal(np.nan_to_num(Ac, nan=-1), np.nan_to_num(Atc, nan=-1))


def test_linearity_and_outlet_sum(oracle, rng):
    from gr2msemidistr_b200 import synth
    d = synth.d8_grid(80, 100).numpy()
    W1 = rng.integers(0, 50, d.shape).astype(float)      # integers: sums are exact, order-free
    W2 = rng.integers(0, 50, d.shape).astype(float)
    A1, A2 = oracle.aread8(d, W1, contcheck=False), oracle.aread8(d, W2, contcheck=False)
    A12 = oracle.aread8(d, 3 * W1 + 2 * W2, contcheck=False)
    ok = ~np.isnan(A12)
    assert np.array_equal(A12[ok], (3 * A1 + 2 * A2)[ok])
    # every valid cell drains to exactly one terminal cell; terminal values sum to the total weight
    pl = oracle.d8_plan(d)
    term = (pl["receiver"] < 0) & (pl["level"] >= 0)
    assert A1.ravel()[term].sum() == W1.ravel()[pl["level"] >= 0].sum()


def test_f32_mode_quantises_like_taudem(oracle, rng):
    from gr2msemidistr_b200 import synth
    d = synth.d8_grid(80, 100).numpy()
    W = rng.uniform(0, 100, d.shape)
    A64 = oracle.aread8(d, W, contcheck=False)
    A32 = oracle.aread8(d, W, f32=True, contcheck=False)
    ok = ~np.isnan(A64)
    assert np.array_equal(A32[ok], A32[ok].astype(np.float32).astype(np.float64))
    rel = np.abs(A32[ok] - A64[ok]) / np.maximum(A64[ok], 1e-30)
    assert 0 < rel.max() < 1e-5


def test_route_month_loop(oracle, rng):
    from gr2msemidistr_b200 import synth
    nrow, ncol = 96, 128
    d = synth.d8_grid(nrow, ncol).numpy()
    src, off, cells = synth.block_polygons(nrow, ncol, 3, 4, margin=2)
    QS = rng.uniform(0, 300, (12, 7))
    QR = oracle.route(d, QS, src, off, cells, contcheck=True, nthreads=2)
    for t in (0, 6):
        W = np.zeros(nrow * ncol); W[src] = QS[:, t]
        A = oracle.aread8(d, W, contcheck=True).ravel()
        for s in range(12):
            v = A[cells[off[s]:off[s + 1]]]
            v = v[~np.isnan(v)]
            assert QR[s, t] == (v.max() if len(v) else -np.inf)
    # duplicates: last assignment wins (R `[<-`), Routing:109-113
    src2 = src.copy(); src2[1] = src2[0]
    QR2 = oracle.route(d, QS, src2, off, cells, contcheck=False)
    W = np.zeros(nrow * ncol); W[src2] = QS[:, 3]
    A = oracle.aread8(d, W, contcheck=False).ravel()
    assert QR2[0, 3] == np.nanmax(A[cells[off[0]:off[1]]])
