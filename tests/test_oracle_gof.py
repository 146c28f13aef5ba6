"""hydroGOF KGE/NSE/rmse restatement (call site R/Optim_GR2MSemiDistr.R:199-216)."""
import numpy as np

from oracle import twin


def test_identities(oracle, rng):
    o = rng.uniform(10, 500, 120)
    n, crit, of = oracle.gof(o, o)
    assert n == 120 and crit[0] == 1.0 and crit[1] == 1.0 and crit[2] == 0.0
    assert np.array_equal(of, [0.0, 0.0, 0.0])
    # NSE(mean(obs), obs) = 0 ; KGE undefined there (sd(sim) = 0 -> cor NaN)
    _, crit, _ = oracle.gof(np.full(120, o.mean()), o)
    assert abs(crit[1]) < 1e-12
    assert np.isnan(crit[0])


def test_five_element_hand_example(oracle):
    sim = np.array([1.0, 2.0, 3.0, 4.0, 6.0])
    obs = np.array([1.0, 3.0, 2.0, 5.0, 5.0])
    _, crit, of = oracle.gof(sim, obs)
    mo = obs.mean()
    nse = 1 - ((obs - sim) ** 2).sum() / ((obs - mo) ** 2).sum()          # 1 - 4/12.8
    rmse = np.sqrt(((sim - obs) ** 2).mean())                             # sqrt(0.8)
    r = np.corrcoef(sim, obs)[0, 1]
    kge = 1 - np.sqrt((r - 1) ** 2 + (sim.std(ddof=1) / obs.std(ddof=1) - 1) ** 2 + (sim.mean() / mo - 1) ** 2)
    assert abs(crit[1] - 0.6875) < 1e-15 and abs(crit[1] - nse) < 1e-15
    assert abs(crit[2] - rmse) < 1e-15 and abs(crit[0] - kge) < 1e-14
    assert np.array_equal(of, [1 - oracle.r_round(kge, 3), 1 - 0.688, 0.894])


def test_warmup_and_na_pairs(oracle, rng):
    sim = rng.uniform(10, 500, 264)
    obs = sim * rng.lognormal(0, 0.2, 264)
    obs[[3, 40, 41, 100, 263]] = np.nan
    n, crit, of = oracle.gof(sim, obs, warmup=36)
    keep = np.arange(264) >= 36
    keep &= ~np.isnan(obs)
    assert n == keep.sum() == 264 - 36 - 4
    n2, crit2, of2 = oracle.gof(sim[keep], obs[keep])
    assert n2 == n and np.array_equal(crit, crit2) and np.array_equal(of, of2)


def test_matches_twin(oracle, rng):
    for _ in range(5):
        sim = rng.uniform(0, 800, 228)
        obs = np.round(sim * rng.lognormal(0, 0.3, 228), 2)
        obs[rng.random(228) < 0.05] = np.nan
        _, crit, of = oracle.gof(sim, obs, warmup=12)
        tc, to = twin.gof(sim, obs, warmup=12)
        assert np.allclose(crit, tc, rtol=1e-14, atol=0)
        assert np.array_equal(of, to)


def test_objective_batch_equals_run_plus_gof(oracle, rng):
    from gr2msemidistr_b200 import synth
    P, E, area, nd = synth.forcing(13, 96)
    par = synth.param_sets(6)
    region = np.zeros(13, np.int32)
    base = oracle.run(P, E, area, nd, region, 1, par[0])
    qobs = synth.qobs_from(oracle.outlet_optim(base["QS"]))
    res = oracle.objective_batch(P, E, area, nd, region, 1, par, qobs, warmup=12, nthreads=3)
    for s in range(6):
        r = oracle.run(P, E, area, nd, region, 1, par[s])
        sink = oracle.outlet_optim(r["QS"])
        assert np.array_equal(res["qt"][s], sink)
        _, crit, of = oracle.gof(sink, qobs, warmup=12)
        assert np.array_equal(res["crit"][s], crit) and np.array_equal(res["of"][s], of)
    # single-subbasin branch: sink is the unrounded QS (Optim:187-188)
    res1 = oracle.objective_batch(P[:1], E[:1], area[:1], nd, region[:1], 1, par, qobs, warmup=12)
    r1 = oracle.run(P[:1], E[:1], area[:1], nd, region[:1], 1, par[2])
    assert np.array_equal(res1["qt"][2], r1["QS"][0])


def test_degenerate_cases_follow_hydrogof(oracle):
    """hydroGOF on degenerate series (R/Optim_GR2MSemiDistr.R:209-212 calls KGE/NSE/rmse on what na.omit
    leaves): cor() is NA when either standard deviation is 0, sd() is NA for one value, NSE is NA when
    obs is constant, everything is NA with no pairs."""
    nan = np.nan
    obs = np.array([3.0, 5.0, 4.0, 9.0, 1.0])
    # constant simulated series: sd(sim) = 0 -> cor NA -> KGE NA; NSE and rmse stay finite
    n, crit, of = oracle.gof(np.full(5, 2.0), obs)
    assert n == 5 and np.isnan(crit[0]) and np.isfinite(crit[1]) and np.isfinite(crit[2])
    assert np.isnan(of[0]) and np.isfinite(of[1]) and np.isfinite(of[2])
    # constant observed series: cor NA and the NSE denominator is 0 -> both NA; rmse finite
    n, crit, _ = oracle.gof(obs, np.full(5, 2.0))
    assert n == 5 and np.isnan(crit[0]) and np.isnan(crit[1]) and abs(crit[2] - np.sqrt(np.mean((obs - 2.0) ** 2))) < 1e-15
    # all-zero observations: mean 0 and sd 0 -> the KGE guard itself says NA
    _, crit, _ = oracle.gof(obs, np.zeros(5))
    assert np.isnan(crit[0]) and np.isnan(crit[1])
    # one valid pair: sd() of one value is NA, NSE denominator 0; rmse = |d|
    n, crit, _ = oracle.gof(np.array([1.0, 2.0, 3.0]), np.array([nan, 2.5, nan]))
    assert n == 1 and np.isnan(crit[0]) and np.isnan(crit[1]) and crit[2] == 0.5
    # no valid pair at all
    n, crit, of = oracle.gof(np.array([1.0, 2.0]), np.array([nan, nan]))
    assert n == 0 and np.all(np.isnan(crit)) and np.all(np.isnan(of))
    # NaN simulation (NaN parameters): rows dropped by na.omit like NA observations
    n, crit, _ = oracle.gof(np.full(4, nan), obs[:4])
    assert n == 0 and np.all(np.isnan(crit))
    # twin agrees on which criteria are defined
    for sim, o in ((np.full(5, 2.0), obs), (obs, np.full(5, 2.0)), (obs, np.zeros(5))):
        tc, _ = twin.gof(sim, o)
        assert np.array_equal(np.isnan(tc), np.isnan(oracle.gof(sim, o)[1]))
