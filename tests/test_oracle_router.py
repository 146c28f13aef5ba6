"""The month-invariant routing operator restated on the CPU (oracle/router_twin.py) against the dense oracle:
the collapse of Routing_GR2MSemiDistr's month loop is exact, bit for bit, whatever the geometry."""
import numpy as np
import pytest

import oracle
from oracle import router_twin
from gr2msemidistr_b200 import synth


def same(a, b):
    return np.array_equal(np.nan_to_num(a, nan=-7.0, neginf=-9.0), np.nan_to_num(b, nan=-7.0, neginf=-9.0))


def geometry(rng, d8, nsub, per_poly):
    nrow, ncol = d8.shape
    valid = np.flatnonzero((d8.ravel() >= 1) & (d8.ravel() <= 8))
    src = rng.choice(valid, nsub, replace=True).astype(np.int32)
    off, cells = np.zeros(nsub + 1, np.int32), []
    for s in range(nsub):
        k = int(rng.integers(0, per_poly + 1))
        r0, c0 = divmod(int(src[s]), ncol)
        rr = np.clip(r0 + rng.integers(-5, 6, k), 0, nrow - 1)
        cc = np.clip(c0 + rng.integers(-5, 6, k), 0, ncol - 1)
        cells.append((rr * ncol + cc).astype(np.int32))
        off[s + 1] = off[s] + k
    return src, off, np.concatenate(cells).astype(np.int32)


@pytest.mark.parametrize("seed,nrow,ncol,nsub", [(1, 30, 40, 6), (2, 60, 80, 40), (3, 90, 70, 150), (4, 25, 25, 1)])
def test_operator_equals_dense_month_loop(seed, nrow, ncol, nsub):
    rng = np.random.default_rng(seed)
    d8 = synth.d8_grid(nrow, ncol, seed=seed).numpy()
    src, off, cells = geometry(rng, d8, nsub, 40)
    QS = rng.uniform(0, 300, (nsub, 7))
    QS[rng.random(QS.shape) < 0.1] = 0.0
    for contcheck in (True, False):
        op = router_twin.build(d8, src, off, cells, contcheck=contcheck)
        assert len(op["nodes"]) <= 2 * nsub
        for f32 in (False, True):
            assert same(router_twin.apply(op, QS, f32=f32),
                        oracle.route(d8, QS, src, off, cells, f32=f32, contcheck=contcheck))


def test_operator_hand_grid():
    N = -1
    d8 = np.array([[1, 1, 8, N, N, N],
                   [1, 1, 1, 8, 1, 1],
                   [1, 1, 1, 1, 1, 1],
                   [1, 1, 2, 1, 1, 1],
                   [N, 1, 1, 1, 1, 1]], np.int16)
    ncol = 6
    src = np.array([0, 3 * ncol, 3 * ncol, 4, 2 * ncol + 5], np.int32)       # a duplicate and one on nodata
    off = np.array([0, 3, 3, 8, 9, 11], np.int32)                          # polygon 1 is empty
    cells = np.array([0, 1, 2, 13, 14, 15, 16, 17, 4, 20, 26], np.int32)
    QS = np.arange(1.0, 21.0).reshape(5, 4)
    op = router_twin.build(d8, src, off, cells)
    got = router_twin.apply(op, QS)
    assert same(got, oracle.route(d8, QS, src, off, cells))
    assert np.all(np.isneginf(got[1]))
    # sources 0 (row 0) and 2 (row 3, the later duplicate) meet where the two branches join
    assert op["node_src"].count(-1) >= 1 and 1 not in op["node_src"] and 3 not in op["node_src"]
