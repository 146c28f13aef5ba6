"""Known-answer and property tests for the GR2M restatement."""
import math

import numpy as np

from oracle import twin


def _forcing(rng, ncell, ntime):
    P = np.round(rng.gamma(0.9, 65.0, (ncell, ntime)), 1)
    E = np.round(np.clip(122 + 40 * np.sin(2 * np.pi * np.arange(ntime) / 12)[None, :] + rng.normal(0, 10, (ncell, ntime)), 0, 210), 1)
    return P, E


def test_hand_trace_three_months(oracle):
    # independent hand evaluation of App. A with math.*, airGR defaults S0 = 0.3*X1, R0 = 0.5*X2
    X1, X2 = 350.0, 0.9
    P, E = [120.0, 0.0, 35.5], [80.0, 110.0, 0.0]
    S, R = 0.3 * X1, 0.5 * X2
    exp_ae, exp_s, exp_q = [], [], []
    for p, e in zip(P, E):
        ws = min(p / X1, 13.0); ev = math.exp(ws); T = (ev - 1 / ev) / (ev + 1 / ev)
        S1 = (S + X1 * T) / (1 + S / X1 * T); P1 = p + S - S1
        ws = min(e / X1, 13.0); ev = math.exp(ws); T = (ev - 1 / ev) / (ev + 1 / ev)
        S2 = S1 * (1 - T) / (1 + (1 - S1 / X1) * T)
        exp_ae.append(S1 - S2)
        sr = S2 / X1; sr = sr * sr * sr + 1
        S = S2 / sr ** (1.0 / 3.0)
        R2 = X2 * (R + P1 + (S2 - S)); Q = R2 * R2 / (R2 + 60); R = R2 - Q
        exp_s.append(S); exp_q.append(Q)
    o = oracle.gr2m_cell(P, E, X1, X2)
    assert np.allclose(o["AE"], exp_ae, rtol=1e-15, atol=0)
    assert np.allclose(o["Prod"], exp_s, rtol=1e-15, atol=0)
    assert np.allclose(o["Qsim"], exp_q, rtol=1e-15, atol=0)
    assert np.allclose(o["state_end"], [S, R], rtol=1e-15, atol=0)
    # first month, coarse sanity: wet month on a 30 % full store produces some runoff
    assert 0 < o["Qsim"][0] < 120 and 0 < o["AE"][0] <= 80


def test_zero_forcing_closed_form(oracle):
    # P = E = 0  =>  S' = S / (1 + (S/X1)^3)^(1/3), Q only from the routing store
    X1, X2, S0, R0 = 500.0, 0.8, 200.0, 30.0
    o = oracle.gr2m_cell([0.0], [0.0], X1, X2, S0, R0)
    S1 = S0 / (1 + (S0 / X1) ** 3) ** (1 / 3)
    R2 = X2 * (R0 + (S0 - S1))
    assert abs(o["Prod"][0] - S1) < 1e-12 and o["AE"][0] == 0.0
    assert abs(o["Qsim"][0] - R2 * R2 / (R2 + 60)) < 1e-12


def test_water_balance_and_bounds(oracle, rng):
    P, E = _forcing(rng, 1, 480)
    for X1, X2 in [(10.976, 0.665), (1000.0, 1.0), (2000.0, 2.0), (1.0, 0.01)]:
        o = oracle.gr2m_cell(P[0], E[0], X1, X2)
        S = np.concatenate([[0.3 * X1], o["Prod"]])
        # reconstruct R from Q: R2 = root of Q = R2^2/(R2+60) => R = R2 - Q, and balance
        R_prev = 0.5 * X2
        for t in range(480):
            Q = o["Qsim"][t]
            R2 = (Q + math.sqrt(Q * Q + 240.0 * Q)) / 2.0
            R1 = R2 / X2
            # P - AE - (S_t - S_{t-1}) = P3 = R1 - R_prev
            lhs = P[0, t] - o["AE"][t] - (S[t + 1] - S[t])
            assert abs(lhs - (R1 - R_prev)) < 1e-9 * max(1.0, abs(lhs))
            R_prev = R2 - Q
            assert 0.0 <= R_prev < 60.0
        assert np.all(o["Prod"] >= 0) and np.all(o["Prod"] <= X1 + 1e-12)
        assert np.all(o["Qsim"] >= 0) and np.all(o["AE"] >= -1e-12)
        assert abs(o["state_end"][1] - R_prev) < 1e-9


def test_clamp_path_and_param_floor(oracle):
    # P/X1 > 13 clamp (documented X1 = 10.976 with P > 143 mm) and X1, X2 floored at 0.01
    o = oracle.gr2m_cell([500.0, 3000.0], [100.0, 100.0], 10.976, 0.665)
    assert np.all(np.isfinite(o["Qsim"])) and o["Qsim"][1] > o["Qsim"][0] > 0
    a = oracle.gr2m_cell([10.0, 20.0], [5.0, 5.0], 0.001, 0.0)
    b = oracle.gr2m_cell([10.0, 20.0], [5.0, 5.0], 0.01, 0.01)
    assert np.array_equal(a["Qsim"], b["Qsim"])


def test_inistate_roundtrip_equals_one_long_run(oracle, rng):
    # EndState -> IniState hand-off (Run:151-179, 224): two half runs == one full run, bit for bit
    P, E = _forcing(rng, 1, 240)
    full = oracle.gr2m_cell(P[0], E[0], 321.0, 0.77)
    a = oracle.gr2m_cell(P[0, :100], E[0, :100], 321.0, 0.77)
    b = oracle.gr2m_cell(P[0, 100:], E[0, 100:], 321.0, 0.77, a["state_end"][0], a["state_end"][1])
    assert np.array_equal(np.concatenate([a["Qsim"], b["Qsim"]]), full["Qsim"])
    assert np.array_equal(b["state_end"], full["state_end"])


def test_tanh_formulations_agree(oracle, rng):
    # parity is unpinned: show the candidate upstream spellings of tanh are equivalent to ~1e-13
    P, E = _forcing(rng, 1, 480)
    base = oracle.gr2m_cell(P[0], E[0], 700.0, 0.9)
    for fl in (oracle.TANH_APPENDIX_A, oracle.TANH_LIBM):
        alt = oracle.gr2m_cell(P[0], E[0], 700.0, 0.9, flags=fl)
        assert np.allclose(alt["Qsim"], base["Qsim"], rtol=1e-12, atol=1e-13)


def test_f32_exponent_mode_is_a_measurable_deviation(oracle, rng):
    # gfortran evaluates the literal (1./3.) in default REAL; effect <= ~7e-9 relative on the store
    P, E = _forcing(rng, 1, 480)
    a = oracle.gr2m_cell(P[0], E[0], 300.0, 0.9)
    b = oracle.gr2m_cell(P[0], E[0], 300.0, 0.9, flags=oracle.PERC_F32_EXPONENT)
    rel = np.max(np.abs(a["Prod"] - b["Prod"]) / a["Prod"])
    assert 1e-12 < rel < 1e-7


def test_run_matches_twin(oracle, rng):
    # forcing correction must agree bit for bit; the recursion to a few ulp (NumPy's SIMD exp/pow
    # are not glibc's)
    ncell, ntime, nreg = 37, 60, 3
    P, E = _forcing(rng, ncell, ntime)
    area = rng.lognormal(np.log(400), 0.6, ncell)
    ndays = np.array([oracle.days_in_month(1981 + t // 12, t % 12 + 1) for t in range(ntime)], np.int32)
    region = rng.integers(0, nreg, ncell).astype(np.int32)
    params = np.concatenate([rng.uniform(1, 2000, nreg), rng.uniform(0.01, 2, nreg),
                             rng.uniform(0.8, 1.2, nreg), rng.uniform(0.8, 1.2, nreg)])
    for f32e in (False, True):
        o = oracle.run(P, E, area, ndays, region, nreg, params, flags=oracle.PERC_F32_EXPONENT if f32e else 0)
        t = twin.run(P, E, area, ndays, region, nreg, params, perc_f32_exponent=f32e)
        assert np.array_equal(o["PR"], t["PR"])
        for k in ("AE", "SM", "RU", "QS", "S_end", "R_end"):
            assert np.allclose(o[k], t[k], rtol=1e-12, atol=1e-11), k
    o2 = oracle.run(P, E, area, ndays, region, nreg, params, nthreads=4)
    o1 = oracle.run(P, E, area, ndays, region, nreg, params, nthreads=1)
    assert all(np.array_equal(o1[k], o2[k]) for k in o1)


def test_days_in_month(oracle):
    assert [oracle.days_in_month(1981, m) for m in range(1, 13)] == [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    assert oracle.days_in_month(1984, 2) == 29 and oracle.days_in_month(2000, 2) == 29 and oracle.days_in_month(1900, 2) == 28


def test_outlet_variants(oracle, rng):
    # Run sums 1-dp rounded QS then rounds to 2 dp (Run:235-236); Optim sums unrounded (Optim:194-195)
    QS = rng.uniform(0, 50, (13, 24))
    ndays = np.full(24, 30, np.int32)
    qt_run = oracle.outlet_run(QS, QS, np.ones(13), ndays)
    assert np.array_equal(qt_run, oracle.r_round(oracle.r_round(QS, 1).sum(0), 2))
    qt_opt = oracle.outlet_optim(QS)
    assert np.allclose(qt_opt, np.round(QS.sum(0), 2), atol=0.0100001)
    assert np.max(np.abs(qt_opt - QS.sum(0))) <= 0.005 + 1e-9
    # nsub == 1 branches (Run:195-199; Optim:187-188)
    ru = rng.uniform(0, 80, (1, 24))
    q1 = oracle.outlet_run(ru, ru, np.array([300.0]), ndays)
    assert np.array_equal(q1, oracle.r_round(300.0 * oracle.r_round(ru[0], 1) / (86.4 * 30), 1))
    assert np.array_equal(oracle.outlet_optim(ru), ru[0])
