"""NumPy twin of oracle/gr2m_oracle.c -- TEST INFRASTRUCTURE ONLY (parity unpinned, see the C header).

An independent, deliberately differently-structured restatement of the same upstream algorithms
(vectorised over cells instead of looping per cell; memoised recursion instead of a Kahn queue).
tests/test_oracle_twin.py requires the two to agree bit-for-bit, so a transcription slip in either
shows up.  Citations are to /root/reference (call sites) and to SURVEY.md appendices A-D (upstream).
"""
from __future__ import annotations

import sys

import numpy as np


def r_round(x, digits: int):
    """R >= 4.0 round(x, digits) for digits >= 1 (call sites Run:101-102 etc.)."""
    x = np.asarray(x, dtype=np.float64)
    a = np.abs(x)
    p10 = np.float64(10.0 ** digits)
    x10 = p10 * a
    i10 = np.floor(x10)
    xd = i10 / p10
    xu = np.ceil(x10) / p10
    du = xu - a
    dd = a - xd
    even = np.mod(i10, 2.0) == 0.0
    r = np.where((dd < du) | ((dd == du) & even), xd, xu)
    with np.errstate(divide="ignore", invalid="ignore"):
        e10 = np.floor(np.log10(a)) + 1
    keep = ~np.isfinite(x) | (x == 0) | (digits + e10 > 15)
    return np.where(keep, x, np.copysign(r, x))


def gr2m(P, E, X1, X2, S0=None, R0=None, perc_f32_exponent=False):
    """airGR GR2M (call site Run:181-184) vectorised over cells.
    P, E: [ncell, ntime] corrected forcing; X1, X2: [ncell]. Returns AE, Prod, Qsim, S_end, R_end."""
    P = np.asarray(P, np.float64)
    E = np.asarray(E, np.float64)
    X1 = np.maximum(np.asarray(X1, np.float64), 0.01)
    X2 = np.maximum(np.asarray(X2, np.float64), 0.01)
    ncell, ntime = P.shape
    S = 0.3 * X1 if S0 is None else np.asarray(S0, np.float64).copy()
    R = 0.5 * X2 if R0 is None else np.asarray(R0, np.float64).copy()
    pexp = np.float64(np.float32(1.0) / np.float32(3.0)) if perc_f32_exponent else 1.0 / 3.0
    AE = np.empty_like(P)
    SM = np.empty_like(P)
    Q = np.empty_like(P)

    def th(w):
        e = np.exp(w)
        return (e - 1.0 / e) / (e + 1.0 / e)

    for t in range(ntime):
        T = th(np.minimum(P[:, t] / X1, 13.0))
        S1 = (S + X1 * T) / (1.0 + (S / X1) * T)
        P1 = P[:, t] + S - S1
        T = th(np.minimum(E[:, t] / X1, 13.0))
        S2 = S1 * (1.0 - T) / (1.0 + (1.0 - S1 / X1) * T)
        AE[:, t] = S1 - S2
        Sr = S2 / X1
        Sr = Sr * Sr * Sr + 1.0
        S = S2 / np.power(Sr, pexp)
        P3 = P1 + (S2 - S)
        R2 = X2 * (R + P3)
        q = R2 * R2 / (R2 + 60.0)
        R = R2 - q
        SM[:, t] = S
        Q[:, t] = q
    return AE, SM, Q, S, R


def run(P, E, area, ndays, region, nreg, params, S0=None, R0=None, perc_f32_exponent=False):
    """Run_GR2MSemiDistr core (Run:82-107,133-187,229). Arrays [ncell, ntime]; outputs unrounded."""
    params = np.asarray(params, np.float64)
    region = np.asarray(region)
    X1, X2 = params[region], params[nreg + region]
    fp, fe = params[2 * nreg + region], params[3 * nreg + region]
    Pc = r_round(fp[:, None] * np.asarray(P, np.float64), 1)
    Ec = r_round(fe[:, None] * np.asarray(E, np.float64), 1)
    AE, SM, Q, S, R = gr2m(Pc, Ec, X1, X2, S0, R0, perc_f32_exponent)
    QS = (np.asarray(area, np.float64)[:, None] * Q) / (86.4 * np.asarray(ndays, np.float64)[None, :])
    return dict(PR=Pc, AE=AE, SM=SM, RU=Q, QS=QS, S_end=S, R_end=R)


def _ld(x):
    return np.asarray(x, dtype=np.longdouble)


def _r_mean(x):
    x = _ld(x)
    s = np.longdouble(0)
    for v in x:
        s += v
    s /= len(x)
    t = np.longdouble(0)
    for v in x:
        t += v - s
    return s + t / len(x)


def gof(sim, obs, warmup=0):
    """hydroGOF KGE/NSE/rmse as called at Optim:199-212. Returns crit[3], of[3]."""
    sim = np.asarray(sim, np.float64)[warmup:]
    obs = np.asarray(obs, np.float64)[warmup:]
    ok = ~(np.isnan(sim) | np.isnan(obs))
    sim, obs = sim[ok], obs[ok]
    n = len(sim)
    ms, mo = _r_mean(sim), _r_mean(obs)
    mod = np.float64(mo)
    sse = np.longdouble(0)
    den = np.longdouble(0)
    for s, o in zip(sim, obs):
        d = np.float64(o - s)
        sse += np.longdouble(d * d)
        a = np.float64(o - mod)
        den += np.longdouble(a * a)
    nse = np.nan if np.float64(den) == 0 else 1.0 - np.float64(sse) / np.float64(den)
    rmse = np.sqrt(np.float64(_r_mean((sim - obs) ** 2)))
    vs = np.longdouble(0)
    vo = np.longdouble(0)
    cv = np.longdouble(0)
    for s, o in zip(sim, obs):
        a, b = np.longdouble(s) - ms, np.longdouble(o) - mo
        vs += a * a
        vo += b * b
        cv += a * b
    sd_s = np.sqrt(np.float64(vs / (n - 1)))
    sd_o = np.sqrt(np.float64(vo / (n - 1)))
    r = np.float64(min(max(cv / (np.sqrt(vs) * np.sqrt(vo)), -1), 1))
    alpha, beta = sd_s / sd_o, np.float64(ms) / mod
    kge = 1.0 - np.sqrt((r - 1.0) ** 2 + (alpha - 1.0) ** 2 + (beta - 1.0) ** 2)
    crit = np.array([kge, nse, rmse])
    of = np.array([1.0 - r_round(kge, 3), 1.0 - r_round(nse, 3), r_round(rmse, 3)])
    return crit, of


_DX = (0, 1, 1, 0, -1, -1, -1, 0, 1)
_DY = (0, 0, -1, -1, -1, 0, 1, 1, 1)


def aread8(flowdir, W, contcheck=True, f32=False):
    """TauDEM AreaD8 -wg by memoised recursion in k = 1..8 order (Routing:117).
    Also returns level[] (0 for heads, 1+max upstream) so the plan's levels can be cross-checked."""
    d = np.asarray(flowdir)
    nrow, ncol = d.shape
    W = np.asarray(W, np.float64).reshape(nrow, ncol)
    A = np.full((nrow, ncol), np.nan)
    lev = np.full((nrow, ncol), -1, np.int32)
    state = np.zeros((nrow, ncol), np.int8)  # 0 new, 1 in progress, 2 done
    acc_t = np.float32 if f32 else np.float64
    sys.setrecursionlimit(max(10000, 4 * (nrow + ncol) + 1000))

    def valid(r, c):
        return 0 <= r < nrow and 0 <= c < ncol and 1 <= d[r, c] <= 8

    def ev(r, c):
        # returns (value or None for nodata, level); None level when on/behind a cycle
        if state[r, c] == 2:
            return (None if np.isnan(A[r, c]) else acc_t(A[r, c])), lev[r, c]
        if state[r, c] == 1:
            raise RecursionError("cycle")
        state[r, c] = 1
        acc = acc_t(W[r, c])
        con = False
        L = 0
        for k in range(1, 9):
            r2, c2 = r + _DY[k], c + _DX[k]
            if not valid(r2, c2):
                con = True
                continue
            if abs(int(d[r2, c2]) - k) == 4:
                v, l2 = ev(r2, c2)
                L = max(L, l2 + 1)
                if v is None:
                    con = True
                else:
                    acc = acc_t(acc + v)
        state[r, c] = 2
        lev[r, c] = L
        A[r, c] = np.nan if (con and contcheck) else np.float64(acc)
        return (None if (con and contcheck) else acc), L

    for r in range(nrow):
        for c in range(ncol):
            if valid(r, c) and state[r, c] == 0:
                ev(r, c)
    return A, lev
