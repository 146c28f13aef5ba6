"""Routed objective inside calibration (north_star: the simulate + route + objective loop).

gr2m_objective_routed_batch composes, for many parameter sets at once, what the reference does with three calls:
Run_GR2MSemiDistr (QS rounded to 0.1, R/Run_GR2MSemiDistr.R:229-235) -> Routing_GR2MSemiDistr (AreaD8 of the weights,
per-polygon maximum, rounded to 0.1, R/Routing_GR2MSemiDistr.R:100-131) -> hydroGOF on the gauge polygon's series
(R/Optim_GR2MSemiDistr.R:199-212).  Parity: the same composition through the CPU oracle (run, r_round, route, gof)."""
import os

import numpy as np
import pytest

from gr2msemidistr_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    assert m.device_count() >= 1
    m.set_device(0)
    return m


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0), initial=0.0))


def oracle_chain(oracle, P, E, area, nd, region, nreg, par, d8, src, off, cells, gauge, qobs, warmup, rounded):
    qt, crit, of = [], [], []
    for p in par:
        sim = oracle.run(P, E, area, nd, region, nreg, p)
        qs = oracle.r_round(sim["QS"], 1) if rounded else sim["QS"]
        qr = oracle.route(d8, qs, src, off, cells, nthreads=8)[gauge]
        if rounded:
            qr = oracle.r_round(qr, 1)
        _, c, o = oracle.gof(qr, qobs, warmup)
        qt.append(qr); crit.append(c); of.append(o)
    return np.array(qt), np.array(crit), np.array(of)


def test_bundled_example_routed_calibration(capi, oracle):
    inp = np.load(os.path.join(G, "c1_inputs.npz"))
    n = 264                                                # 01/1981 .. 12/2002, WarmUp 36 (Optim example)
    P, E, area, nd = inp["P"][:, :n], inp["E"][:, :n], inp["area"], inp["ndays"][:n]
    region = np.zeros(13, np.int32)
    par = synth.param_sets(10)
    qobs = inp["qobs"][:n]
    d8, src, off, cells = inp["flowdir"], inp["src_cell"], inp["poly_off"], inp["poly_cells"]
    with capi.Plan(d8) as plan, plan.router(src, off, cells) as router, capi.Basin(P, E, area, nd, region, 1) as b:
        full = plan.route(oracle.r_round(oracle.run(P, E, area, nd, region, 1, par[0])["QS"], 1), src, off, cells)
        gauge = int(np.argmax(np.nanmean(np.where(np.isfinite(full), full, np.nan), axis=1)))    # the polygon the basin drains through
        for rounded in (True, False):
            fl = 0 if rounded else capi.SINK_UNROUNDED
            want_qt, want_crit, want_of = oracle_chain(oracle, P, E, area, nd, region, 1, par, d8, src, off, cells, gauge, qobs, 36, rounded)
            got = {w: b.objective_routed_batch(router, par, qobs, gauge, 36, w, fl, want_qt=True) for w in (0, 1, 2)}
            if rounded:      # rounded series: identical except where a value sits on a rounding tie
                d = np.abs(got[0]["qt"] - want_qt)
                assert np.all(d <= 0.1 + 1e-9) and np.mean(d > 1e-9) <= 1e-4
                same = np.all(d <= 1e-9, axis=1)
                assert same.mean() > 0.9
                assert relerr(got[0]["crit"][same], want_crit[same]) <= TOL
                for w in (0, 1, 2):
                    assert np.array_equal(got[w]["of"][same], want_of[same, w])
            else:
                assert relerr(got[0]["qt"], want_qt) <= TOL
                assert relerr(got[0]["crit"], want_crit) <= TOL
        # the polygon every subbasin drains through accumulates the whole basin: its unrounded routed series is the outlet
        # sum Optim scores (Optim:186-196), up to the order of the additions; an upstream polygon sees less
        plain = b.objective_batch(par, qobs, 36, capi.OF_KGE, capi.SINK_UNROUNDED, want_qt=True)
        assert np.isfinite(got[0]["of"]).all() and relerr(got[0]["qt"], plain["qt"]) <= 1e-12
        other = (gauge + 1) % 13
        up = b.objective_routed_batch(router, par[:4], qobs, other, 36, capi.OF_KGE, capi.SINK_UNROUNDED, want_qt=True)
        assert np.all(up["qt"] <= got[0]["qt"][:4] + 1e-9) and np.any(up["qt"] < got[0]["qt"][:4] - 1e-6)


def test_synthetic_many_sets_chunks_and_regions(capi, oracle, rng, monkeypatch):
    # several regions (general recursion kernel), more subbasins than a tile, sets processed in several chunks
    nrow, ncol = 96, 128
    d8 = synth.d8_grid(nrow, ncol, seed=5).numpy()
    src, off, cells = synth.block_polygons(nrow, ncol, 6, 8, margin=1)
    nsub, ntime, nreg = len(src), 40, 3
    P, E, area, nd = synth.forcing(nsub, ntime, seed=6)
    region = rng.integers(0, nreg, nsub).astype(np.int32)
    par = synth.param_sets(37, nreg, seed=7)
    with capi.Plan(d8) as plan, plan.router(src, off, cells) as router, capi.Basin(P, E, area, nd, region, nreg) as b:
        base = oracle.route(d8, oracle.run(P, E, area, nd, region, nreg, par[0])["QS"], src, off, cells)
        gauge = int(np.nanargmax(np.where(np.isfinite(base), base, -1).mean(axis=1)))
        qobs = synth.qobs_from(base[gauge])
        want_qt, want_crit, _ = oracle_chain(oracle, P, E, area, nd, region, nreg, par, d8, src, off, cells, gauge, qobs, 5, False)
        got = b.objective_routed_batch(router, par, qobs, gauge, 5, capi.OF_NSE, capi.SINK_UNROUNDED, want_qt=True)
        assert relerr(got["qt"], want_qt) <= TOL and relerr(got["crit"], want_crit) <= TOL
        one = b.objective_routed_batch(router, par[5:6], qobs, gauge, 5, capi.OF_NSE, capi.SINK_UNROUNDED, want_qt=True)
        assert np.array_equal(one["qt"][0], got["qt"][5])          # a set's result does not depend on its batch
    with pytest.raises(capi.GR2MError):
        with capi.Plan(d8) as plan, plan.router(src[:5], off[:6], cells[:off[5]]) as router, capi.Basin(P, E, area, nd, region, nreg) as b:
            b.objective_routed_batch(router, par, qobs, 0)         # router built for another set of subbasins


def test_fused_gauge_path_matches_the_materialised_one(capi, oracle, rng, monkeypatch):
    # one region: the routed objective runs fused (weighted, per-subbasin-rounded outlet sums of the gauge's maximal
    # candidate nodes, no QS / QR planes).  Gauges with one and with several maximal candidates, rounded and unrounded,
    # against the materialised path (GR2M_ROUTED_MATERIALISE=1) and the oracle chain.
    nrow, ncol = 160, 192
    d8 = synth.d8_grid(nrow, ncol, seed=15).numpy()
    src, off, cells = synth.block_polygons(nrow, ncol, 8, 8, margin=1)
    nsub, ntime = len(src), 48
    P, E, area, nd = synth.forcing(nsub, ntime, seed=16)
    region = np.zeros(nsub, np.int32)
    par = synth.param_sets(21, seed=17)
    with capi.Plan(d8) as plan, plan.router(src, off, cells) as router, capi.Basin(P, E, area, nd, region, 1) as b:
        base = oracle.route(d8, oracle.run(P, E, area, nd, region, 1, par[0])["QS"], src, off, cells)
        mean = np.where(np.isfinite(base), base, -1).mean(axis=1)
        order = np.argsort(-mean)
        qobs = synth.qobs_from(base[order[0]])
        # the upstream sets behind the fused path: routing ones gives, at every polygon, the size of its largest set
        ones = router.apply(np.ones((nsub, 1)))[:, 0]
        for g in range(0, nsub, 5):
            k, owner = router.gauge_upstream(g)
            sizes = np.bincount(owner[owner >= 0], minlength=max(k, 1))
            assert (k == 0 and not ones[g] > 0) or (k > 0 and sizes.max() == ones[g] and len(sizes) == k and sizes.min() > 0)
        checked = fused = 0
        # square blocks on a tilted plane are crossed by many parallel streams: a polygon with more than four maximal
        # candidates keeps the materialised path, so try a spread of polygons and require that some ran fused
        for gauge in [int(g) for g in order[::max(1, len(order) // 12)]]:
            for fl in (0, capi.SINK_UNROUNDED):
                n0 = capi.kernel_launches()
                got = b.objective_routed_batch(router, par, qobs, gauge, 5, capi.OF_KGE, fl, want_qt=True)
                fused += capi.kernel_launches() - n0 <= 6        # <= 4 recursion launches + combine + objective
                monkeypatch.setenv("GR2M_ROUTED_MATERIALISE", "1")
                mat = b.objective_routed_batch(router, par, qobs, gauge, 5, capi.OF_KGE, fl, want_qt=True)
                monkeypatch.delenv("GR2M_ROUTED_MATERIALISE")
                if fl:
                    assert relerr(got["qt"], mat["qt"]) <= 1e-12
                    want_qt, want_crit, _ = oracle_chain(oracle, P, E, area, nd, region, 1, par[:6], d8, src, off, cells, gauge, qobs, 5, False)
                    assert relerr(got["qt"][:6], want_qt) <= TOL
                    ok = np.isfinite(want_crit)
                    assert relerr(got["crit"][:6][ok], want_crit[ok]) <= TOL
                else:       # sums of values rounded to 0.1, rounded to 0.1: the same numbers unless a sum sits on a tie
                    d = np.abs(got["qt"] - mat["qt"])
                    assert np.all(d <= 0.1 + 1e-9) and np.mean(d > 1e-9) <= 1e-4
                checked += 1
        assert checked >= 16 and fused >= 4, (checked, fused)
