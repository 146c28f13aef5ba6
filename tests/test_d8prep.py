"""CPU reference of the D8 preprocessing rule (oracle/d8prep.py) -- properties and hand cases."""
import os

import numpy as np

from oracle import d8prep

N = -32768
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def terrain(rng, nrow, ncol, quant=1.0, holes=0.0):
    z = rng.normal(0, 1, (nrow, ncol))
    for _ in range(3):
        z = (z + np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) / 5
    z = np.round(z * 40 / quant) * quant          # integer-valued: many flats and pits
    if holes:
        z[rng.random((nrow, ncol)) < holes] = np.nan
    return z


def test_hand_cases():
    # a pit in a bowl: filled to the spill elevation, drains over the lowest rim cell
    z = np.array([[9, 9, 9, 9, 9], [9, 5, 5, 5, 9], [9, 5, 1, 5, 9], [9, 5, 5, 5, 7], [9, 9, 9, 9, 9]], float)
    filled, fd = d8prep.fill_and_flowdir(z)
    assert filled[2, 2] == 7 and filled[1, 1] == 7 and filled[3, 4] == 7
    assert fd[3, 4] == 1                                   # the rim cell drains off-grid to the east (first invalid k)
    assert fd[3, 3] == 1 and fd[2, 2] in (8, 1, 7)         # flat points one BFS step closer to the spill
    assert (fd[1:4, 1:4] > 0).all()
    # a tilted plane: everything flows east, the last column leaves the grid
    z = np.tile(np.arange(6, 0, -1.0), (4, 1))
    _, fd = d8prep.fill_and_flowdir(z)
    assert (fd[1:-1, :-1] == 1).all() and (fd[:, -1] > 0).all()
    # nodata cells stay nodata and act as outlets for their neighbours
    z = np.full((5, 5), 3.0); z[2, 2] = np.nan
    _, fd = d8prep.fill_and_flowdir(z)
    assert fd[2, 2] == N and fd[1, 1] == 8 and fd[2, 1] == 1


def test_every_cell_drains_without_cycles(oracle, rng):
    for shape, holes in (((60, 80), 0.0), ((90, 70), 0.03)):
        z = terrain(rng, *shape, holes=holes)
        filled, fd = d8prep.fill_and_flowdir(z)
        valid = np.isfinite(z)
        assert np.array_equal(fd > 0, valid)
        assert np.all(filled[valid] >= z[valid]) and (filled[valid] > z[valid]).sum() > 0
        pl = oracle.d8_plan(fd)
        assert len(pl["order"]) == valid.sum()                # Kahn reaches every cell: no cycles
        # flow never goes uphill on the filled surface
        rec = pl["receiver"]
        has = rec >= 0
        f = filled.ravel()
        assert np.all(f[rec[has]] <= f[np.flatnonzero(has)])


def test_fixture_flowdir_is_this_rule():
    inp = np.load(os.path.join(G, "c1_inputs.npz"))
    z = np.where(inp["dem"] == N, np.nan, inp["dem"].astype(float))
    _, fd = d8prep.fill_and_flowdir(z)
    assert np.array_equal(fd, inp["flowdir"])
