"""Device depression fill + D8 directions vs the CPU reference rule: bit-exact."""
import os

import numpy as np
import pytest

from oracle import d8prep
from test_d8prep import terrain

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    assert m.device_count() >= 1
    m.set_device(0)
    return m


def _check(capi, z):
    valid = np.isfinite(z)
    filled, fd = d8prep.fill_and_flowdir(z)
    got_fd, got_f = capi.flowdir_from_dem(np.where(valid, z, 0.0), valid, want_filled=True)
    assert np.array_equal(np.nan_to_num(got_f, nan=-9e9), np.nan_to_num(filled, nan=-9e9))
    assert np.array_equal(got_fd, fd)


def test_bundled_dem(capi):
    inp = np.load(os.path.join(G, "c1_inputs.npz"))
    z = np.where(inp["dem"] == -32768, np.nan, inp["dem"].astype(float))
    _check(capi, z)
    valid = np.isfinite(z)
    assert np.array_equal(capi.flowdir_from_dem(np.where(valid, z, 0.0), valid), inp["flowdir"])


@pytest.mark.parametrize("shape,quant,holes", [((33, 65), 1.0, 0.0), ((200, 300), 1.0, 0.02), ((257, 129), 5.0, 0.0),
                                               ((96, 96), 0.01, 0.05)])
def test_synthetic_terrain_with_pits_and_flats(capi, rng, shape, quant, holes):
    _check(capi, terrain(rng, *shape, quant=quant, holes=holes))


def test_edge_cases(capi):
    _check(capi, np.full((7, 9), 3.0))                       # one big flat: every cell one BFS ring further in
    _check(capi, np.full((4, 4), np.nan))                    # nothing valid
    z = np.full((5, 5), np.nan); z[2, 2] = 1.0
    _check(capi, z)                                          # a single cell
    z = np.add.outer(np.arange(40.0), np.arange(50.0)); z[10:30, 10:40] = 5.0
    _check(capi, z)                                          # a plateau cut into a slope (a pit to fill)
    _check(capi, np.arange(12.0)[None, :])                   # one row
