"""GPU parity: D8 plan (bit-exact integers) and Weighted Flow Accumulation vs the CPU oracle."""
import numpy as np
import pytest

from gr2msemidistr_b200 import synth

pytestmark = pytest.mark.gpu
N = -32768
PLAN_KEYS = ("receiver", "inmask", "indeg", "level", "order", "level_off", "contam")


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    assert m.device_count() >= 1, "gpu tests need a CUDA device"
    m.set_device(0)
    return m


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0), initial=0.0))


@pytest.mark.parametrize("shape", [(5, 5), (64, 96), (251, 429), (700, 900)])
def test_plan_bit_exact(capi, oracle, shape):
    d = synth.d8_grid(*shape, seed=shape[0]).numpy()
    ref = oracle.d8_plan(d)
    with capi.Plan(d) as p:
        got = p.export()
    assert got["nlevels"] == ref["nlevels"]
    for k in PLAN_KEYS:
        assert np.array_equal(got[k], ref[k]), k


def test_plan_edge_cases(capi, oracle, rng):
    # all nodata; a single valid cell; random directions with cycles (cycle cells and everything
    # behind them are never evaluated, as in TauDEM); one row; one column
    cases = [np.full((4, 6), N, np.int16), np.array([[N, N, N], [N, 1, N], [N, N, N]], np.int16),
             rng.integers(0, 10, (40, 50)).astype(np.int16), rng.integers(1, 9, (1, 64)).astype(np.int16),
             rng.integers(1, 9, (64, 1)).astype(np.int16)]
    for d in cases:
        ref = oracle.d8_plan(d)
        with capi.Plan(d) as p:
            got = p.export()
            for k in PLAN_KEYS:
                assert np.array_equal(got[k], ref[k]), k
            W = rng.uniform(0, 10, (3,) + d.shape)
            for fl in (0, capi.WFA_NO_CONTAMINATION, capi.WFA_DATAFLOW, capi.WFA_LEVEL_SYNC, capi.WFA_BY_BASIN, capi.WFA_NO_OVERLAP):
                A = p.accumulate(W, fl)
                Ar = np.stack([oracle.aread8(d, W[m], contcheck=not (fl & capi.WFA_NO_CONTAMINATION)) for m in range(3)])
                assert np.array_equal(np.isnan(A), np.isnan(Ar))
                assert relerr(np.nan_to_num(A), np.nan_to_num(Ar)) <= 1e-10


@pytest.mark.parametrize("nmonth", [1, 5, 12, 33, 70])
def test_accumulate_dense_matches_oracle(capi, oracle, rng, nmonth):
    d = synth.d8_grid(150, 210, seed=7).numpy()
    W = rng.uniform(0, 100, (nmonth, 150, 210))
    with capi.Plan(d) as p:
        for fl in (0, capi.WFA_DATAFLOW, capi.WFA_LEVEL_SYNC, capi.WFA_BY_BASIN, capi.WFA_NO_OVERLAP, capi.WFA_NO_CONTAMINATION):
            A = p.accumulate(W, fl)
            cc = not (fl & capi.WFA_NO_CONTAMINATION)
            for m in {0, nmonth // 2, nmonth - 1}:
                Ar = oracle.aread8(d, W[m], contcheck=cc)
                assert np.array_equal(np.isnan(A[m]), np.isnan(Ar))
                # same k-ordered FP64 additions as the oracle => identical bits
                assert np.array_equal(np.nan_to_num(A[m]), np.nan_to_num(Ar))
        # TauDEM-faithful float32 mode is bit-exact against the oracle's float32 mode
        A32 = p.accumulate(W, capi.WFA_F32)
        for m in {0, nmonth - 1}:
            Ar = oracle.aread8(d, W[m], f32=True, contcheck=True)
            assert np.array_equal(np.nan_to_num(A32[m], nan=-1), np.nan_to_num(Ar, nan=-1))


def test_dataflow_tail_on_long_thin_network(capi, oracle, rng):
    # one long meandering channel: thousands of levels with 1-2 cells each (the routing tail)
    nrow, ncol = 64, 257
    d = np.full((nrow, ncol), N, np.int16)
    for r in range(1, nrow - 1, 2):
        if (r // 2) % 2 == 0:
            d[r, 1:ncol - 1] = 1; d[r, ncol - 2] = 7; d[r + 1, ncol - 2] = 7
        else:
            d[r, 1:ncol - 1] = 5; d[r, 1] = 7; d[r + 1, 1] = 7
    d[nrow - 2:, :] = N
    ref = oracle.d8_plan(d)
    assert ref["nlevels"] > 5000
    W = rng.uniform(0, 10, (8, nrow, ncol))
    with capi.Plan(d) as p:
        got = p.export()
        for k in PLAN_KEYS:
            assert np.array_equal(got[k], ref[k]), k
        for fl in (capi.WFA_NO_CONTAMINATION, capi.WFA_NO_CONTAMINATION | capi.WFA_DATAFLOW):
            A = p.accumulate(W, fl)
            for m in (0, 7):
                Ar = oracle.aread8(d, W[m], contcheck=False)
                assert np.array_equal(np.nan_to_num(A[m], nan=-1), np.nan_to_num(Ar, nan=-1))


def test_route_matches_oracle_c1_shape(capi, oracle, rng):
    # config C1 shape: 251 x 429 grid, 13 subbasins, 432 months
    nrow, ncol = 251, 429
    d = synth.d8_grid(nrow, ncol, seed=13).numpy()
    src, off, cells = synth.block_polygons(nrow, ncol, 3, 5, margin=3)
    src, off = src[:13], off[:14]
    cells = cells[:off[-1]]
    QS = rng.uniform(0, 400, (13, 432))
    ref = oracle.route(d, QS, src, off, cells, nthreads=8)
    ref32 = oracle.route(d, QS, src, off, cells, f32=True, nthreads=8)
    with capi.Plan(d) as p:
        QR = p.route(QS, src, off, cells)
        QR_ls = p.route(QS, src, off, cells, capi.WFA_DATAFLOW)
        QR32 = p.route(QS, src, off, cells, capi.WFA_F32)
        QRnc = p.route(QS, src, off, cells, capi.WFA_NO_CONTAMINATION)
    assert np.array_equal(QR, ref) and np.array_equal(QR_ls, ref)
    assert np.array_equal(QR32, ref32)
    assert np.array_equal(QRnc, oracle.route(d, QS, src, off, cells, contcheck=False, nthreads=8))


def test_route_month_chunking_is_invisible(capi, rng, monkeypatch):
    # big grids are routed in month chunks to bound the dense buffer; the result must not change
    nrow, ncol = 251, 429
    d = synth.d8_grid(nrow, ncol, seed=13).numpy()
    src, off, cells = synth.block_polygons(nrow, ncol, 3, 5, margin=3)
    QS = rng.uniform(0, 400, (15, 100))
    with capi.Plan(d) as p:
        whole = p.route(QS, src, off, cells)
        monkeypatch.setenv("WFA_ROUTE_BUDGET_MB", "30")          # 107 679 cells x 8 B -> chunks of 32 months
        chunked = p.route(QS, src, off, cells)
        chunked32 = p.route(QS, src, off, cells, capi.WFA_F32)
        monkeypatch.delenv("WFA_ROUTE_BUDGET_MB")
        whole32 = p.route(QS, src, off, cells, capi.WFA_F32)
    assert np.array_equal(whole, chunked) and np.array_equal(whole32, chunked32)


def test_route_edge_cases(capi, oracle, rng):
    nrow, ncol = 96, 128
    d = synth.d8_grid(nrow, ncol).numpy()
    src, off, cells = synth.block_polygons(nrow, ncol, 3, 4, margin=2)
    QS = rng.uniform(0, 300, (12, 7))
    with capi.Plan(d) as p:
        # duplicate source cells: the last assignment wins (R `[<-`)
        src2 = src.copy(); src2[1] = src2[0]; src2[5] = src2[0]
        assert np.array_equal(p.route(QS, src2, off, cells, capi.WFA_NO_CONTAMINATION),
                              oracle.route(d, QS, src2, off, cells, contcheck=False))
        # an empty polygon gives -Inf (R max() of nothing), a single month works (Update mode)
        off2 = off.copy(); off2[4:] -= off2[4] - off2[3]
        cells2 = np.concatenate([cells[:off[3]], cells[off[4]:]])
        got = p.route(QS[:, :1], src, off2, cells2)
        assert np.array_equal(got, oracle.route(d, QS[:, :1], src, off2, cells2))
        assert got[3, 0] == -np.inf
        # out-of-range cells are rejected, not read
        bad = src.copy(); bad[0] = nrow * ncol
        with pytest.raises(capi.GR2MError):
            p.route(QS, bad, off, cells)


def test_per_basin_sweep_is_used_and_bit_identical(capi, oracle, rng):
    # a grid that splits into many outlet basins: WFA_BY_BASIN regroups the upper levels by basin;
    # it must give the same bits as the grid-wide level sweep and as the oracle
    nrow, ncol = 900, 1200
    d = synth.d8_grid(nrow, ncol, seed=5).numpy()
    W = rng.uniform(0, 100, (33, nrow, ncol))
    NC, BB = capi.WFA_NO_CONTAMINATION, capi.WFA_BY_BASIN
    with capi.Plan(d) as p:
        info = p.basin_info()
        assert info["enabled"] and info["nbasins"] >= 32 and info["nupper"] > 0
        A = p.accumulate(W, NC | BB)
        B = p.accumulate(W, NC | capi.WFA_LEVEL_SYNC)
        # 32 and 20 months take the shared-memory-forwarding kernel (nmonth <= lanes per cell), 33 the general one
        C32 = p.accumulate(W[:32], NC | BB)
        D32 = p.accumulate(W[:32], NC | BB | capi.WFA_NO_FORWARD)
        C20 = p.accumulate(W[:20], NC | BB | capi.WFA_F32)
        D20 = p.accumulate(W[:20], NC | capi.WFA_F32 | capi.WFA_LEVEL_SYNC)
    assert np.array_equal(np.nan_to_num(A, nan=-1), np.nan_to_num(B, nan=-1))
    assert np.array_equal(np.nan_to_num(C32, nan=-1), np.nan_to_num(A[:32], nan=-1))
    assert np.array_equal(np.nan_to_num(D32, nan=-1), np.nan_to_num(A[:32], nan=-1))
    assert np.array_equal(np.nan_to_num(C20, nan=-1), np.nan_to_num(D20, nan=-1))
    for m in (0, 32):
        Ar = oracle.aread8(d, W[m], contcheck=False)
        assert np.array_equal(np.nan_to_num(A[m], nan=-1), np.nan_to_num(Ar, nan=-1))


@pytest.mark.parametrize("width", [0, 64])
def test_reach_sweep_is_used_and_bit_identical(capi, oracle, rng, width, monkeypatch):
    # default sweep of the upper network: one lane group per river reach, reaches in dependency levels.
    # Same k-ordered additions as every other sweep => same bits.  width=64 moves the cut far down so
    # that the reach network has many confluences and short reaches.
    if width:
        monkeypatch.setenv("WFA_COMP_WIDTH", str(width))
    nrow, ncol = 700, 1100
    d = synth.d8_grid(nrow, ncol, seed=11).numpy()
    NC = capi.WFA_NO_CONTAMINATION
    with capi.Plan(d) as p:
        ri = p.reach_info()
        assert ri["enabled"] and ri["nreaches"] > 0 and 0 < ri["nreach_levels"] < p.nlevels
        for nm, fl in ((33, 0), (32, 0), (7, 0), (1, 0), (20, capi.WFA_F32), (40, capi.WFA_F32)):
            W = rng.uniform(0, 100, (nm, nrow, ncol))
            A = p.accumulate(W, NC | fl)
            B = p.accumulate(W, NC | fl | capi.WFA_LEVEL_SYNC)
            assert np.array_equal(np.nan_to_num(A, nan=-1), np.nan_to_num(B, nan=-1)), (nm, fl)
        W = rng.uniform(0, 100, (3, nrow, ncol))
        A = p.accumulate(W, 0)
        for m in range(3):
            Ar = oracle.aread8(d, W[m], contcheck=True)
            assert np.array_equal(np.nan_to_num(A[m], nan=-1), np.nan_to_num(Ar, nan=-1))


def test_linearity_and_mass_conservation_large(capi, rng):
    # BASELINE-scale property test (oracle-free): integer weights make sums exact, so
    # route(3 W1 + 2 W2) == 3 route(W1) + 2 route(W2) bit for bit, and terminal cells hold the total
    nrow, ncol, nm = 1500, 2000, 4
    d = synth.d8_grid(nrow, ncol, seed=99).numpy()
    W1 = rng.integers(0, 50, (nm, nrow, ncol)).astype(float)
    W2 = rng.integers(0, 50, (nm, nrow, ncol)).astype(float)
    with capi.Plan(d) as p:
        pl = p.export()
        fl = capi.WFA_NO_CONTAMINATION
        A1, A2, A12 = p.accumulate(W1, fl), p.accumulate(W2, fl), p.accumulate(3 * W1 + 2 * W2, fl)
    ok = ~np.isnan(A12)
    assert np.array_equal(A12[ok], (3 * A1 + 2 * A2)[ok])
    term = ((pl["receiver"] < 0) & (pl["level"] >= 0)).reshape(nrow, ncol)
    valid = (pl["level"] >= 0).reshape(nrow, ncol)
    for m in range(nm):
        assert A1[m][term].sum() == W1[m][valid].sum()


@pytest.mark.parametrize("terrain,width", [("tilted", 0), ("tilted", 64), ("fractal", 256), ("fractal", 0)])
def test_chain_segments_are_used_and_bit_identical(capi, oracle, rng, terrain, width, monkeypatch):
    # default sweep below the cut: a lane group computes a head cell and carries the running sum through the cells
    # that have no other non-source inflow (csrc/wfa.cu, k_accum_segments).  Same k-ordered additions => same bits
    # as the cell-by-cell level sweep, as grid-wide level sync and as the CPU oracle; FP64 and float32, any month count.
    import torch
    if width:
        monkeypatch.setenv("WFA_COMP_WIDTH", str(width))
    nrow, ncol = 640, 900
    if terrain == "tilted":
        d = synth.d8_grid(nrow, ncol, seed=13).numpy()
    else:      # smooth sheet flow with filled depressions: few sources, long single-inflow chains, flats
        z, valid = synth.fractal_dem(nrow, ncol, relief=4.0)
        d = capi.flowdir_from_dem(z.numpy(), valid.numpy())
    NC = capi.WFA_NO_CONTAMINATION
    with capi.Plan(d) as p:
        si = p.segment_info()
        if width:                    # many levels below the cut: the segments must really absorb cells
            assert si["enabled"] and si["levels"] > 9 and si["ncells"] > 1.15 * si["nsegments"] > 0
        for nm, fl in ((33, 0), (32, 0), (7, 0), (1, 0), (70, 0), (20, capi.WFA_F32), (40, capi.WFA_F32)):
            W = rng.uniform(0, 100, (nm, nrow, ncol))
            A = p.accumulate(W, NC | fl)
            B = p.accumulate(W, NC | fl | capi.WFA_NO_SEGMENTS)
            Cc = p.accumulate(W, NC | fl | capi.WFA_LEVEL_SYNC)
            assert np.array_equal(np.nan_to_num(A, nan=-1), np.nan_to_num(B, nan=-1)), (nm, fl)
            assert np.array_equal(np.nan_to_num(A, nan=-1), np.nan_to_num(Cc, nan=-1)), (nm, fl)
        W = rng.uniform(0, 100, (2, nrow, ncol))
        A = p.accumulate(W, 0)
        A32 = p.accumulate(W, capi.WFA_F32)
        for m in range(2):
            assert np.array_equal(np.nan_to_num(A[m], nan=-1), np.nan_to_num(oracle.aread8(d, W[m], contcheck=True), nan=-1))
            assert np.array_equal(np.nan_to_num(A32[m], nan=-1), np.nan_to_num(oracle.aread8(d, W[m], f32=True, contcheck=True), nan=-1))
