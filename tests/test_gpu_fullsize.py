"""BASELINE.json full-size configurations, checked through size-independent properties (the CPU oracle
would need hours at these sizes)."""
import numpy as np
import pytest

from gr2msemidistr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    assert m.device_count() >= 1
    m.set_device(0)
    return m


def test_c4_full_sweep_is_deterministic_and_batch_independent(capi, oracle):
    """C4: 1e5 subbasins x 1e4 parameter sets x 480 months (4.8e11 cell-months per pass)."""
    ncell, nsets, ntime = 100_000, 10_000, 480
    P, E, area, nd = synth.forcing(ncell, ntime)
    region = np.zeros(ncell, np.int32)
    par = synth.param_sets(nsets)
    qobs = synth.qobs_from(np.full(ntime, 100.0) * (1 + 0.5 * np.sin(np.arange(ntime) / 6.0)))
    pick = np.array([0, 1, 777, 4999, 5000, 9998, 9999])
    with capi.Basin(P, E, area, nd, region, 1) as b:
        full = b.objective_batch(par, qobs, 36, capi.OF_KGE)
        again = b.objective_batch(par, qobs, 36, capi.OF_KGE)
        few = b.objective_batch(par[pick], qobs, 36, capi.OF_KGE, want_qt=True)
        unr = b.objective_batch(par[pick[:2]], qobs, 36, capi.OF_KGE, capi.SINK_UNROUNDED, want_qt=True)
    # run-to-run determinism, and a set's value does not depend on the batch it is evaluated in
    assert np.array_equal(full["of"], again["of"], equal_nan=True)
    assert np.array_equal(full["of"][pick], few["of"], equal_nan=True)
    assert np.array_equal(full["crit"][pick], few["crit"], equal_nan=True)
    assert np.isfinite(full["of"]).all() and np.all(few["qt"] >= 0)
    # anchor two of the 1e4 sets against the oracle on a 2 000-subbasin slice scaled by linearity of the
    # outlet sum: sum over all cells == sum of 50 disjoint slices, here checked on the first slice only
    sl = slice(0, 2000)
    with capi.Basin(P[sl], E[sl], area[sl], nd, region[sl], 1) as b:
        part = b.objective_batch(par[pick[:2]], None, 0, capi.OF_KGE, capi.SINK_UNROUNDED, want_crit=False, want_qt=True)["qt"]
    ref = oracle.objective_batch(P[sl], E[sl], area[sl], nd, region[sl], 1, par[pick[:2]], qobs, 0,
                                 flags=oracle.SINK_UNROUNDED, nthreads=8)["qt"]
    assert np.max(np.abs(part - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-10
    assert np.all(unr["qt"] >= part - 1e-6)      # the full-basin outlet can only exceed one slice of it


def test_c5_full_grid_mass_conservation(capi):
    """C5: 20 000 x 20 000 D8 grid; integer weights make every sum exact, so the accumulation at the
    terminal cells must equal the total weight over the swept cells, month by month."""
    import torch
    n, ld = 20_000, 4
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    d8 = synth.d8_grid(n, n, device=dev)
    with capi.Plan(dev_ptr=d8.data_ptr(), shape=(n, n)) as p:
        del d8
        p.set_stream(st.cuda_stream)
        assert p.nlevels > 5000 and p.norder > 0.9 * n * n
        info = p.basin_info()
        assert info["enabled"] and info["nbasins"] > 100
        pl = p.export()
        valid = torch.from_numpy(pl["level"] >= 0).to(dev)
        term = torch.from_numpy((pl["receiver"] < 0) & (pl["level"] >= 0)).to(dev)
        indeg_total = int(pl["indeg"].astype(np.int64).sum())
        assert indeg_total == int(((pl["receiver"] >= 0)).sum())          # every edge counted once
        del pl
        g = torch.Generator(device=dev).manual_seed(5)
        A = torch.randint(0, 64, (n * n, ld), generator=g, device=dev, dtype=torch.int32).to(torch.float64)
        total = (A * valid[:, None]).sum(0)
        p.accumulate_dev(A.data_ptr(), ld, ld, capi.WFA_NO_CONTAMINATION)
        st.synchronize()
        got = (A * term[:, None]).sum(0)
        assert torch.equal(got, total)
        B = A.clone()
        # the grid-wide level sweep gives the same bits as the per-basin sweep
        A.copy_(torch.randint(0, 64, (n * n, ld), generator=torch.Generator(device=dev).manual_seed(5), device=dev,
                              dtype=torch.int32).to(torch.float64))
        p.accumulate_dev(A.data_ptr(), ld, ld, capi.WFA_NO_CONTAMINATION | capi.WFA_LEVEL_SYNC)
        st.synchronize()
        assert torch.equal(A, B)
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
