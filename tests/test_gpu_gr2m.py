"""GPU parity: recursion + objective through the C ABI vs the CPU oracle (run with -m gpu on a B200).

Tolerance (north_star: <= 1e-10 relative, FP64): |gpu - ref| <= 1e-10 * max(|ref|, 1) in the
output's unit (mm or m3/s); the forcing correction round(f*x,1) must be bit-exact.
"""
import numpy as np
import pytest

from gr2msemidistr_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-10


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0), initial=0.0))


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    assert m.device_count() >= 1, "gpu tests need a CUDA device"
    m.set_device(0)
    return m


def test_device_math_building_blocks(capi, oracle, rng):
    # R round(x,1): fast path and exact path bit-identical to the oracle, incl. exact decimal ties
    x = np.concatenate([rng.uniform(0, 4000, 200000), 1.186 * (rng.integers(0, 40000, 200000) / 10.0),
                        rng.integers(0, 80000, 100000) / 20.0, [0.0, 0.05, 0.15, 0.25, 0.35, 29.65, 1e-9, 3839.95, -0.15, -7.25]])
    # adversarial: 10x within 1e-12 ... 1e-3 of a decimal tie, at magnitudes up to where ulp(10x) nears the
    # 1e-5 window of the cheap path (it must hand those to the exact path, not decide them)
    k = np.concatenate([rng.integers(0, 400000, 20000), rng.integers(10**6, 10**9, 20000),
                        rng.integers(10**9, 2**33, 5000)]).astype(np.float64)
    delta = rng.choice([0.0, 1e-12, 1e-10, 1e-8, 1e-6, 5e-6, 9e-6, 1.1e-5, 2e-5, 1e-4, 1e-3], k.size) * rng.choice([-1, 1], k.size)
    x = np.concatenate([x, (k + 0.5 + delta) / 10.0, -(k[:5000] + 0.5 + delta[:5000]) / 10.0])
    want = oracle.r_round(x, 1)
    assert np.array_equal(capi.debug_math(0, x), want)
    assert np.array_equal(capi.debug_math(1, x), want)
    for d in (2, 3):
        v = np.concatenate([rng.uniform(-2, 5000, 100000), rng.integers(0, 10 ** 6, 50000) / 2000.0])
        assert np.array_equal(capi.debug_math(5, v, np.full(v.size, float(d))), oracle.r_round(v, d))
    # exp on [0, 26], division, reciprocal cube root on [1, 2]: a few ulp
    e = rng.uniform(0, 26, 200000)
    assert np.max(np.abs(capi.debug_math(2, e) / np.exp(e) - 1)) < 7e-16      # <= 2 ulp + NumPy's own 1
    n, dd = rng.uniform(0, 5000, 200000), rng.uniform(2, 4e11, 200000)
    assert np.max(np.abs(capi.debug_math(3, n, dd) / (n / dd) - 1)) < 4.5e-16
    assert np.max(np.abs(capi.debug_math(6, n, dd) / (n / dd) - 1)) < 4.5e-16      # the one-Newton-step variant in use
    assert np.max(np.abs(capi.debug_math(7, n, dd) / (n / dd) - 1)) < 6e-16        # cubic-step variant (not in use unless it passes here)
    assert np.max(np.abs(capi.debug_math(13, n, dd) / (n / dd) - 1)) < 6e-16       # same with the numerator multiplied in early (shipped)
    assert np.max(np.abs(capi.debug_math(8, e) / np.exp(e) - 1)) < 7e-16           # 256-entry table, degree 4
    # 2048-entry table, degree 3, one-FMA reduction (shipped calibration step): <= 1.4e-15 at x = 26 by construction
    assert np.max(np.abs(capi.debug_math(10, e) / np.exp(e) - 1)) < 2.5e-15
    small = e[e < 3]
    assert np.max(np.abs(capi.debug_math(10, small) / np.exp(small) - 1)) < 7e-16
    # the shipped form: argument in table-index units, exp(x)/2 (x * 2048/ln2 costs <= 2 ulp of x on the exponent)
    assert np.max(np.abs(capi.debug_math(11, e) / np.exp(e) - 1)) < 8e-15
    assert np.max(np.abs(capi.debug_math(11, small) / np.exp(small) - 1)) < 1.2e-15
    big = np.array([26.0, 27.0, 1e3, 7.6e5])                  # clamped like min(WS, 13): exp(26) up to 2.2e-5
    assert np.max(np.abs(capi.debug_math(11, big) / np.exp(26.0) - 1)) < 3e-5
    assert np.array_equal(capi.debug_math(9, x), want)                             # rint by the conversion unit
    c = np.concatenate([rng.uniform(1, 2, 200000), [1.0, 2.0, np.nextafter(1.0, 2), np.nextafter(2.0, 1)], 1 + rng.uniform(0, 1, 20000) ** 3])
    assert np.max(np.abs(capi.debug_math(4, c) * np.cbrt(c) - 1)) < 7e-16
    assert np.max(np.abs(capi.debug_math(12, c) * np.cbrt(c) - 1)) < 7e-16         # seed without the XU pipe (c in [1, 2])
    assert np.isnan(capi.debug_math(12, np.array([np.nan])))[0]
    # shipped: piecewise-linear table seed (c in [1, 2], 2.0 included: it has its own table entry)
    cc = np.concatenate([c[(c >= 1) & (c <= 2)], np.linspace(1.0, 2.0, 100001), 1 + np.arange(257) / 256.0,
                         np.nextafter(1 + np.arange(1, 257) / 256.0, 0.0)])
    assert np.max(np.abs(capi.debug_math(14, cc) * np.cbrt(cc) - 1)) < 7e-16


@pytest.mark.parametrize("flags", [0, 16, 1, 17])
def test_run_c1_shape_matches_oracle(capi, oracle, flags):
    # config C1 shape: 13 subbasins x 432 months x 1 set, the documented parameter set
    P, E, area, nd = synth.forcing(13, 432)
    region = np.zeros(13, np.int32)
    par = np.array(synth.SET0)
    ref = oracle.run(P, E, area, nd, region, 1, par, flags=flags & 1)
    with capi.Basin(P, E, area, nd, region, 1) as b:
        got = b.run(par, flags=flags)
    assert np.array_equal(got["PR"], ref["PR"])
    for k in ("AE", "SM", "RU", "QS", "S_end", "R_end"):
        assert relerr(got[k], ref[k]) <= TOL, k


def test_run_multi_region_ragged_and_inistate(capi, oracle, rng):
    # 3 regions, ncell not a multiple of 32, ntime not a multiple of the 16-month stage
    ncell, ntime, nreg = 77, 61, 3
    P, E, area, nd = synth.forcing(ncell, ntime, seed=5)
    region = rng.integers(0, nreg, ncell).astype(np.int32)
    par = synth.param_sets(4, nreg, seed=9)[3]
    ref = oracle.run(P, E, area, nd, region, nreg, par)
    got = capi.run_once(P, E, area, nd, region, nreg, par)
    assert np.array_equal(got["PR"], ref["PR"])
    for k in ("AE", "SM", "RU", "QS", "S_end", "R_end"):
        assert relerr(got[k], ref[k]) <= TOL, k
    # EndState -> IniState hand-off (Run:151-179): two half runs equal the long run
    h = 30
    a = capi.run_once(P[:, :h], E[:, :h], area, nd[:h], region, nreg, par)
    b = capi.run_once(P[:, h:], E[:, h:], area, nd[h:], region, nreg, par, S0=a["S_end"], R0=a["R_end"])
    assert relerr(np.concatenate([a["QS"], b["QS"]], axis=1), ref["QS"]) <= TOL
    assert relerr(b["S_end"], ref["S_end"]) <= TOL and relerr(b["R_end"], ref["R_end"]) <= TOL
    # single month (Update mode, Run:138-142 needs a dummy month in airGR; not here)
    one = capi.run_once(P[:, :1], E[:, :1], area, nd[:1], region, nreg, par)
    assert relerr(one["QS"], ref["QS"][:, :1]) <= TOL


def test_run_c3_shape(capi, oracle):
    # config C3: 3 600 subbasins x 480 months, 14 regional parameter sets
    ncell, ntime, nreg = 3600, 480, 14
    P, E, area, nd = synth.forcing(ncell, ntime, seed=3)
    region = (np.arange(ncell) * 7 % nreg).astype(np.int32)
    par = synth.param_sets(2, nreg, seed=4)[1]
    ref = oracle.run(P, E, area, nd, region, nreg, par, nthreads=8)
    with capi.Basin(P, E, area, nd, region, nreg) as b:
        got = b.run(par)
    assert np.array_equal(got["PR"], ref["PR"])
    for k in ("AE", "SM", "RU", "QS"):
        assert relerr(got[k], ref[k]) <= TOL, k


def test_extreme_parameters_and_forcing(capi, oracle):
    # bounds of the calibration box, X1/X2 below the airGR floor, P/X1 > 13 clamp, zero forcing
    P, E, area, nd = synth.forcing(40, 48, seed=11)
    P[0] = 0.0; E[1] = 0.0; P[2] = 3000.0; P[3, ::2] = 0.0
    region = np.zeros(40, np.int32)
    for par in ([1.0, 0.01, 0.8, 0.8], [2000.0, 2.0, 1.2, 1.2], [0.001, 0.0, 1.0, 1.0], [10.976, 0.665, 1.186, 1.169]):
        par = np.array(par)
        ref = oracle.run(P, E, area, nd, region, 1, par)
        got = capi.run_once(P, E, area, nd, region, 1, par)
        assert np.array_equal(got["PR"], ref["PR"])
        for k in ("AE", "SM", "RU", "QS"):
            assert relerr(got[k], ref[k]) <= TOL, (k, par)


def _check_objective(capi, oracle, P, E, area, nd, region, nreg, par, warmup, flags=0):
    base = oracle.run(P, E, area, nd, region, nreg, par[0])
    qobs = synth.qobs_from(oracle.outlet_optim(base["QS"]))
    with capi.Basin(P, E, area, nd, region, nreg) as b:
        unr = b.objective_batch(par, qobs, warmup, capi.OF_NSE, flags | capi.SINK_UNROUNDED, want_qt=True)
        rnd = {w: b.objective_batch(par, qobs, warmup, w, flags, want_qt=True) for w in (0, 1, 2)}
        one = b.objective_batch(par[1:2], qobs, warmup, capi.OF_KGE, flags)      # the OFUN drop-in call
    ref_u = oracle.objective_batch(P, E, area, nd, region, nreg, par, qobs, warmup, flags=(flags & 1) | oracle.SINK_UNROUNDED, nthreads=8)
    ref_r = oracle.objective_batch(P, E, area, nd, region, nreg, par, qobs, warmup, flags=flags & 1, nthreads=8)
    # unrounded sink and raw criteria: strict tolerance
    assert relerr(unr["qt"], ref_u["qt"]) <= TOL
    ok = np.isfinite(ref_u["crit"])
    assert np.array_equal(np.isfinite(unr["crit"]), ok)
    assert relerr(unr["crit"][ok], ref_u["crit"][ok]) <= TOL
    # rounded path (Optim:195, 210-212): identical except where a sum sits on a rounding tie.  A value lands
    # within the 1e-13 parity margin of a tie with probability ~1e-11 (sink, quantum 0.01) / ~1e-10 (criterion,
    # quantum 0.001); measured: 0 of 2.64 M sink values and 0 of 30 000 criteria (profiles/r1_parity_margins.jsonl),
    # so the bound is "at most one quantum, at most 1e-5 / 1e-4 of the values"
    dq = np.abs(rnd[1]["qt"] - ref_r["qt"])
    assert np.all(dq <= 0.01 + 1e-9) and np.mean(dq > 1e-9) <= 1e-5
    for w in (0, 1, 2):
        d = np.abs(rnd[w]["of"] - ref_r["of"][:, w])
        okw = np.isfinite(ref_r["of"][:, w])
        assert np.array_equal(np.isfinite(rnd[w]["of"]), okw)
        assert np.all(d[okw] <= 1e-3 + 1e-9) and np.mean(d[okw] > 1e-9) <= 1e-4
    assert abs(one["of"][0] - rnd[0]["of"][1]) == 0.0


def test_objective_c2_shape(capi, oracle):
    # config C2 as quoted: 13 subbasins x 264 months (WarmUp 36) x 10^4 candidate sets in one batch
    P, E, area, nd = synth.forcing(13, 264)
    _check_objective(capi, oracle, P, E, area, nd, np.zeros(13, np.int32), 1, synth.param_sets(10000), 36)


def test_objective_multi_tile_multi_region(capi, oracle, rng):
    # several cell tiles and cell splits, ragged sizes, 2 regions, sets not a multiple of 8
    ncell, ntime, nreg = 333, 75, 2
    P, E, area, nd = synth.forcing(ncell, ntime, seed=21)
    region = rng.integers(0, nreg, ncell).astype(np.int32)
    _check_objective(capi, oracle, P, E, area, nd, region, nreg, synth.param_sets(61, nreg, seed=22), 12)
    _check_objective(capi, oracle, P, E, area, nd, region, nreg, synth.param_sets(9, nreg, seed=23), 0, flags=capi.MATH_LITERAL)
    _check_objective(capi, oracle, P, E, area, nd, region, nreg, synth.param_sets(9, nreg, seed=24), 0, flags=capi.PERC_F32_EXPONENT)


def test_objective_extreme_parameter_sets(capi, oracle):
    # corners of the calibration box, parameters below the airGR floor, the P/X1 > 13 clamp, factors that
    # create exact decimal ties (1.186 * 25.0), zero and huge forcing -- through the calibration kernel
    P, E, area, nd = synth.forcing(70, 96, seed=61)
    P[0] = 0.0; E[1] = 0.0; P[2] = 3000.0; P[3, ::2] = 25.0; P[4] = 75.0
    region = np.zeros(70, np.int32)
    par = np.array([[1.0, 0.01, 0.8, 0.8], [2000.0, 2.0, 1.2, 1.2], [1.0, 2.0, 1.2, 0.8], [2000.0, 0.01, 0.8, 1.2],
                    [0.001, 0.0, 1.0, 1.0], [10.976, 0.665, 1.186, 1.169], [500.0, 1.0, 1.0, 1.0], [50.0, 0.5, 1.15, 0.85]])
    qobs = synth.qobs_from(np.full(96, 200.0))
    for fl in (0, capi.MATH_LITERAL):
        with capi.Basin(P, E, area, nd, region, 1) as b:
            got = b.objective_batch(par, qobs, 12, capi.OF_NSE, fl | capi.SINK_UNROUNDED, want_qt=True)
        ref = oracle.objective_batch(P, E, area, nd, region, 1, par, qobs, 12, flags=oracle.SINK_UNROUNDED, nthreads=4)
        assert relerr(got["qt"], ref["qt"]) <= TOL
        ok = np.isfinite(ref["crit"])
        assert np.array_equal(np.isfinite(got["crit"]), ok) and relerr(got["crit"][ok], ref["crit"][ok]) <= TOL


def test_objective_rounding_ties_in_table_mode(capi, oracle, rng):
    # the table-mode kernel decides round(f x, 1) in fixed point (one FMA, tie window 2^-26 on 10 f x, chunk redo through
    # R's candidate comparison): forcing placed on, just beside and far from decimal ties of f x for "nice" and random
    # factors, zero months (10 f x = 0 exactly), values that are exact eighths of a tenth -- one region, several tile pairs
    ncell, ntime = 150, 64
    P, E, area, nd = synth.forcing(ncell, ntime, seed=71)
    k = rng.integers(0, 4000, (ncell, ntime)).astype(np.float64)
    delta = rng.choice([0.0, 1e-13, -1e-13, 1e-10, -1e-10, 3e-9, -3e-9, 2e-8, -2e-8, 1e-6, 0.0125, 0.025], (ncell, ntime))
    tie = rng.random((ncell, ntime)) < 0.5
    P = np.where(tie, (k + 0.5 + delta) / 10.0, P)                     # f = 1: 10 f x = k + 1/2 + delta
    E = np.where(rng.random((ncell, ntime)) < 0.3, (rng.integers(0, 2000, (ncell, ntime)) + 0.5) / 12.5, E)   # f = 1.25
    P[5] = 0.0; E[6] = 0.0; P[7] = 0.05; P[8, ::3] = 6553.45
    par = synth.param_sets(24, seed=72)
    par[:8, 2:] = [[1.0, 1.0], [1.0, 1.25], [1.25, 1.0], [0.8, 1.2], [1.186, 1.169], [1.001, 0.999], [1.2, 0.8], [1.125, 0.875]]
    par[8:12, 0] = [1.0, 3.0, 40.0, 2000.0]                            # with and without the min(P / X1, 13) clamp
    region = np.zeros(ncell, np.int32)
    qobs = synth.qobs_from(np.full(ntime, 150.0))
    with capi.Basin(P, E, area, nd, region, 1) as b:
        got = b.objective_batch(par, qobs, 6, capi.OF_NSE, capi.SINK_UNROUNDED, want_qt=True)
    ref = oracle.objective_batch(P, E, area, nd, region, 1, par, qobs, 6, flags=oracle.SINK_UNROUNDED, nthreads=4)
    assert relerr(got["qt"], ref["qt"]) <= TOL


def test_objective_single_subbasin_branch(capi, oracle):
    # nsub == 1: sink is the unrounded QS (Optim:187-188)
    P, E, area, nd = synth.forcing(1, 120, seed=31)
    par = synth.param_sets(17)
    base = oracle.run(P, E, area, nd, np.zeros(1, np.int32), 1, par[0])
    qobs = synth.qobs_from(base["QS"][0])
    with capi.Basin(P, E, area, nd, np.zeros(1, np.int32), 1) as b:
        got = b.objective_batch(par, qobs, 12, capi.OF_RMSE, want_qt=True)
    ref = oracle.objective_batch(P, E, area, nd, np.zeros(1, np.int32), 1, par, qobs, 12)
    assert relerr(got["qt"], ref["qt"]) <= TOL
    assert relerr(got["crit"], ref["crit"]) <= TOL


def test_sharding_by_parameter_set_is_bit_identical(capi):
    # multi-GPU design: each rank evaluates a slice of the sets against replicated
    # forcing; a slice must give exactly the bits the full batch gives
    P, E, area, nd = synth.forcing(200, 96, seed=41)
    region = np.zeros(200, np.int32)
    par = synth.param_sets(96)
    qobs = synth.qobs_from(np.full(96, 50.0))
    with capi.Basin(P, E, area, nd, region, 1) as b:
        full = b.objective_batch(par, qobs, 12, capi.OF_NSE, want_qt=True)
        parts = [b.objective_batch(par[i:i + 24], qobs, 12, capi.OF_NSE, want_qt=True) for i in range(0, 96, 24)]
    assert np.array_equal(np.concatenate([p["qt"] for p in parts]), full["qt"])
    assert np.array_equal(np.concatenate([p["of"] for p in parts]), full["of"], equal_nan=True)


def test_large_sweep_properties(capi):
    # BASELINE-scale slice where the oracle is too slow: size-independent properties.
    # (i) the result for a set does not depend on which other sets share the batch (checked above);
    # (ii) permuting the subbasins permutes nothing in the outlet sum beyond re-association (<= 1e-12);
    # (iii) duplicating every subbasin with half the area leaves the outlet series unchanged.
    ncell, ntime, nsets = 4096, 480, 256
    P, E, area, nd = synth.forcing(ncell, ntime, seed=51)
    region = np.zeros(ncell, np.int32)
    par = synth.param_sets(nsets)
    with capi.Basin(P, E, area, nd, region, 1) as b:
        q0 = b.objective_batch(par, None, 0, capi.OF_NSE, capi.SINK_UNROUNDED, want_crit=False, want_qt=True)["qt"]
    perm = np.random.default_rng(1).permutation(ncell)
    with capi.Basin(P[perm], E[perm], area[perm], nd, region, 1) as b:
        q1 = b.objective_batch(par, None, 0, capi.OF_NSE, capi.SINK_UNROUNDED, want_crit=False, want_qt=True)["qt"]
    assert relerr(q1, q0) <= 1e-12
    P2, E2, a2 = np.repeat(P, 2, 0), np.repeat(E, 2, 0), np.repeat(area / 2, 2)
    with capi.Basin(P2, E2, a2, nd, np.zeros(2 * ncell, np.int32), 1) as b:
        q2 = b.objective_batch(par, None, 0, capi.OF_NSE, capi.SINK_UNROUNDED, want_crit=False, want_qt=True)["qt"]
    assert relerr(q2, q0) <= 1e-12
    assert np.all(np.isfinite(q0)) and np.all(q0 >= 0)


def test_objective_degenerate_series_match_oracle(capi, oracle):
    """hydroGOF's undefined cases (R/Optim_GR2MSemiDistr.R:209-212): a criterion that is NA in R (cor() with a
    zero standard deviation, sd() of one value, a zero NSE denominator, no pairs left by na.omit) must be NaN
    on the device too -- the first build clamped a NaN correlation to -1 and returned a finite KGE."""
    ncell, ntime, wu = 40, 72, 12
    P, E, area, nd = synth.forcing(ncell, ntime, seed=71)
    region = np.zeros(ncell, np.int32)
    par = np.vstack([synth.param_sets(13, seed=72), [[np.nan, 0.5, 1.0, 1.0], [300.0, np.nan, 1.0, 1.0], [300.0, 0.5, np.nan, 1.0]]])
    base = oracle.run(P, E, area, nd, region, 1, par[0])
    good = synth.qobs_from(oracle.outlet_optim(base["QS"]))
    one_pair = np.full(ntime, np.nan); one_pair[40] = 17.25
    two_pairs = one_pair.copy(); two_pairs[41] = 3.5
    only_warmup = np.full(ntime, np.nan); only_warmup[:wu] = 5.0
    cases = {
        "regular": (area, good),
        "constant sim (areas so small that round(sum, 2) = 0)": (area * 1e-9, good),
        "constant obs": (area, np.full(ntime, 42.0)),
        "all-zero obs": (area, np.zeros(ntime)),
        "one valid pair": (area, one_pair),
        "two valid pairs": (area, two_pairs),
        "no valid pair": (area, np.full(ntime, np.nan)),
        "valid pairs only inside the warm-up": (area, only_warmup),
    }
    for name, (ar, qobs) in cases.items():
        for fl in (0, capi.SINK_UNROUNDED):
            ref = oracle.objective_batch(P, E, ar, nd, region, 1, par, qobs, wu, flags=oracle.SINK_UNROUNDED if fl else 0, nthreads=4)
            with capi.Basin(P, E, ar, nd, region, 1) as b:
                got = {w: b.objective_batch(par, qobs, wu, w, fl, want_qt=True) for w in (0, 1, 2)}
            nan_ref = np.isnan(ref["crit"])
            assert np.array_equal(np.isnan(got[0]["crit"]), nan_ref), name
            assert np.array_equal(np.isnan(got[0]["qt"]), np.isnan(ref["qt"])), name
            assert relerr(got[0]["crit"][~nan_ref], ref["crit"][~nan_ref]) <= TOL, name
            for w in (0, 1, 2):      # the value SCE-UA minimises: NaN exactly where R's is NA
                assert np.array_equal(np.isnan(got[w]["of"]), np.isnan(ref["of"][:, w])), (name, w)
    # sanity of the cases themselves (so the test cannot pass vacuously)
    ref = oracle.objective_batch(P, E, area * 1e-9, nd, region, 1, par[:2], good, wu)
    assert np.all(ref["qt"] == 0.0) and np.all(np.isnan(ref["crit"][:, 0])) and np.all(np.isfinite(ref["crit"][:, 1:]))
    ref = oracle.objective_batch(P, E, area, nd, region, 1, par[:2], np.full(ntime, 42.0), wu)
    assert np.all(np.isnan(ref["crit"][:, :2])) and np.all(np.isfinite(ref["crit"][:, 2]))


@pytest.mark.parametrize("nreg", [1, 2])
def test_sums_do_not_depend_on_batch_size_or_rank_count(capi, nreg):
    # Outlet sums are accumulated per canonical group of cell tiles (a function of the number of cells only), so the
    # bits of a set's series are the same whether it is evaluated alone, in a batch of 8 or of 640 -- i.e. for any
    # split of the sets over ranks -- although the launch is cut into very different blocks in each case.
    ncell, ntime = 9000, 48
    P, E, area, nd = synth.forcing(ncell, ntime, seed=91)
    region = (np.arange(ncell) % nreg).astype(np.int32)
    par = synth.param_sets(640, nreg, seed=92)
    with capi.Basin(P, E, area, nd, region, nreg) as b:
        full = b.objective_batch(par, None, 0, capi.OF_NSE, capi.SINK_UNROUNDED, want_crit=False, want_qt=True)["qt"]
        for lo, hi in ((0, 1), (5, 13), (100, 260), (633, 640)):
            part = b.objective_batch(par[lo:hi], None, 0, capi.OF_NSE, capi.SINK_UNROUNDED, want_crit=False, want_qt=True)["qt"]
            assert np.array_equal(part, full[lo:hi]), (lo, hi)


def test_regions_grouped_by_subbasin_take_the_table_mode_passes(capi, oracle, rng):
    # several regions whose subbasins are (mostly) consecutive: the table-mode recursion runs once per region with the
    # other regions' areas masked (tile pairs without a cell of the region are skipped); interleaved regions keep the
    # general kernel.  Same parity bar, sums independent of the batch, and the launch count tells which path ran.
    ncell, ntime, nreg = 1500, 60, 3
    P, E, area, nd = synth.forcing(ncell, ntime, seed=81)
    P[7] = 0.0; P[700, ::2] = 25.0; P[1400] = 2500.0
    region = np.repeat(np.arange(nreg), [517, 420, 563]).astype(np.int32)        # ragged boundaries inside tile pairs
    region[[3, 900, 1499]] = [2, 0, 1]                                            # a few strays
    par = synth.param_sets(45, nreg, seed=82)
    par[1, :nreg] = [1.0, 2000.0, 30.0]                                           # clamp in one region only
    base = oracle.run(P, E, area, nd, region, nreg, par[0])
    qobs = synth.qobs_from(oracle.outlet_optim(base["QS"]))
    for fl in (0, capi.PERC_F32_EXPONENT):
        with capi.Basin(P, E, area, nd, region, nreg) as b:
            n0 = capi.kernel_launches()
            got = b.objective_batch(par, qobs, 12, capi.OF_NSE, fl | capi.SINK_UNROUNDED, want_qt=True)
            assert capi.kernel_launches() - n0 == nreg + 2       # parameter gather + one recursion pass per region + objective
            one = b.objective_batch(par[9:10], qobs, 12, capi.OF_NSE, fl | capi.SINK_UNROUNDED, want_qt=True)
        ref = oracle.objective_batch(P, E, area, nd, region, nreg, par, qobs, 12, flags=(fl & 1) | oracle.SINK_UNROUNDED, nthreads=4)
        assert relerr(got["qt"], ref["qt"]) <= TOL
        ok = np.isfinite(ref["crit"])
        assert np.array_equal(np.isfinite(got["crit"]), ok) and relerr(got["crit"][ok], ref["crit"][ok]) <= TOL
        assert np.array_equal(one["qt"][0], got["qt"][9])
    with capi.Basin(P, E, area, nd, (np.arange(ncell) % nreg).astype(np.int32), nreg) as b:      # interleaved: general kernel
        n0 = capi.kernel_launches()
        b.objective_batch(par, qobs, 12, capi.OF_NSE, capi.SINK_UNROUNDED)
        assert capi.kernel_launches() - n0 == 2
