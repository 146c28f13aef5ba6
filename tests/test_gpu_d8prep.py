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


def test_device_pointer_variant_and_fractal_terrain(capi):
    # the survey-spec terrain of bench.py --route-terrain filled, small: device-resident call == host call == CPU rule
    import torch
    from gr2msemidistr_b200 import synth
    dev = torch.device("cuda", 0)
    z, valid = synth.fractal_dem(150, 210, device=dev, relief=10.0)      # steep enough for many depressions on a small grid
    v8 = valid.to(torch.uint8).contiguous()
    fd = torch.empty((150, 210), dtype=torch.int16, device=dev)
    filled = torch.empty((150, 210), dtype=torch.float64, device=dev)
    capi.flowdir_from_dem_dev(z.data_ptr(), v8.data_ptr(), 150, 210, fd.data_ptr(), filled.data_ptr())
    zh, vh = z.cpu().numpy(), valid.cpu().numpy()
    want_f, want_fd = d8prep.fill_and_flowdir(np.where(vh, zh, np.nan))
    assert np.array_equal(fd.cpu().numpy(), want_fd)
    assert np.array_equal(np.nan_to_num(filled.cpu().numpy(), nan=-9e9), np.nan_to_num(want_f, nan=-9e9))
    assert np.array_equal(capi.flowdir_from_dem(np.where(vh, zh, 0.0), vh), want_fd)
    assert np.any((want_f > zh + 1e-9)[vh])                    # there was something to fill
    fd2 = torch.empty_like(fd)
    capi.flowdir_from_dem_dev(z.data_ptr(), v8.data_ptr(), 150, 210, fd2.data_ptr())      # without the filled surface
    assert torch.equal(fd, fd2)
