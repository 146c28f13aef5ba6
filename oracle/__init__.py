"""CPU oracle for the GR2MSemiDistr hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference ships no tests or golden vectors and its arithmetic lives in
third-party code (airGR, TauDEM, hydroGOF, R's round) that cannot run in this image; see the
header of ``gr2m_oracle.c``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  Nothing under
``gr2msemidistr_b200/`` does.

``build()`` compiles ``gr2m_oracle.c`` with gcc into ``oracle/_build/``; the functions below are thin
ctypes wrappers returning NumPy arrays laid out like the R matrices (column-major ``[ntime x ncell]``,
exposed here as C-contiguous ``[ncell, ntime]`` arrays -- the same bytes).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgr2m_oracle.so")
_lib = None

PERC_F32_EXPONENT = 1
TANH_APPENDIX_A = 2
TANH_LIBM = 4
SINK_UNROUNDED = 8


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gr2m_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_r_round.restype = C.c_double
        _lib.orc_r_round.argtypes = [C.c_double, C.c_int]
        _lib.orc_days_in_month.restype = C.c_int
        _lib.orc_days_in_month.argtypes = [C.c_int, C.c_int]
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def r_round(x, digits: int):
    scalar = np.ndim(x) == 0
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    out = np.empty_like(x)
    lib().orc_r_round_vec(_p(x, C.c_double), C.c_int(x.size), C.c_int(digits), _p(out, C.c_double))
    return float(out.reshape(-1)[0]) if scalar else out


def days_in_month(year: int, month: int) -> int:
    return lib().orc_days_in_month(year, month)


def gr2m_cell(P, E, X1, X2, S0=None, R0=None, flags=0):
    """airGR RunModel_GR2M on already-corrected forcing. Returns dict(AE, Prod, Qsim, state_end)."""
    P, E = _f64(P), _f64(E)
    n = P.size
    AE, SM, Q, st = np.empty(n), np.empty(n), np.empty(n), np.empty(2)
    has = S0 is not None and R0 is not None
    lib().orc_gr2m_cell(C.c_int(n), _p(P, C.c_double), _p(E, C.c_double), C.c_double(X1), C.c_double(X2),
                        C.c_int(int(has)), C.c_double(S0 if has else 0.0), C.c_double(R0 if has else 0.0),
                        C.c_int(flags), _p(AE, C.c_double), _p(SM, C.c_double), _p(Q, C.c_double),
                        _p(st, C.c_double))
    return dict(AE=AE, Prod=SM, Qsim=Q, state_end=st)


def run(P, E, area, ndays, region, nreg, params, S0=None, R0=None, flags=0, nthreads=1):
    """Run_GR2MSemiDistr simulation core. P, E: [ncell, ntime] (== R column-major [ntime x ncell]).
    Returns dict of UNROUNDED PR, AE, SM, RU, QS [ncell, ntime] and S_end, R_end [ncell]."""
    P, E, area, params = _f64(P), _f64(E), _f64(area), _f64(params)
    ndays, region = _i32(ndays), _i32(region)
    S0, R0 = _f64(S0), _f64(R0)
    ncell, ntime = P.shape
    out = {k: np.empty((ncell, ntime)) for k in ("PR", "AE", "SM", "RU", "QS")}
    out["S_end"], out["R_end"] = np.empty(ncell), np.empty(ncell)
    rc = lib().orc_run(C.c_int(ntime), C.c_int(ncell), _p(P, C.c_double), _p(E, C.c_double),
                       _p(area, C.c_double), _p(ndays, C.c_int), _p(region, C.c_int), C.c_int(nreg),
                       _p(params, C.c_double), _p(S0, C.c_double), _p(R0, C.c_double), C.c_int(flags),
                       C.c_int(nthreads), *[_p(out[k], C.c_double) for k in ("PR", "AE", "SM", "RU", "QS", "S_end", "R_end")])
    if rc != 0:
        raise ValueError(f"orc_run failed: {rc}")
    return out


def outlet_run(QS, RU, area, ndays):
    QS, RU, area, ndays = _f64(QS), _f64(RU), _f64(area), _i32(ndays)
    ncell, ntime = QS.shape
    qt = np.empty(ntime)
    lib().orc_outlet_run(C.c_int(ntime), C.c_int(ncell), _p(QS, C.c_double), _p(RU, C.c_double),
                         _p(area, C.c_double), _p(ndays, C.c_int), _p(qt, C.c_double))
    return qt


def outlet_optim(QS, unrounded=False):
    QS = _f64(QS)
    ncell, ntime = QS.shape
    qt = np.empty(ntime)
    lib().orc_outlet_optim(C.c_int(ntime), C.c_int(ncell), _p(QS, C.c_double), _p(qt, C.c_double),
                           C.c_int(int(unrounded)))
    return qt


def gof(sim, obs, warmup=0):
    """Returns (n_valid, crit[KGE,NSE,RMSE], of[1-round(KGE,3), 1-round(NSE,3), round(RMSE,3)])."""
    sim, obs = _f64(sim), _f64(obs)
    crit, of = np.empty(3), np.empty(3)
    n = lib().orc_gof(C.c_int(sim.size), _p(sim, C.c_double), _p(obs, C.c_double), C.c_int(warmup or 0),
                      _p(crit, C.c_double), _p(of, C.c_double))
    return n, crit, of


def objective_batch(P, E, area, ndays, region, nreg, params, qobs, warmup=0, flags=0, nthreads=1,
                    want_qt=True):
    """OFUN of Optim_GR2MSemiDistr for many sets. params: [nsets, 4*nreg].
    Returns dict(crit [nsets,3], of [nsets,3], qt [nsets, ntime])."""
    P, E, area, params = _f64(P), _f64(E), _f64(area), _f64(params)
    ndays, region, qobs = _i32(ndays), _i32(region), _f64(qobs)
    ncell, ntime = P.shape
    nsets = params.shape[0]
    crit, of = np.full((nsets, 3), np.nan), np.full((nsets, 3), np.nan)
    qt = np.empty((nsets, ntime)) if want_qt else None
    rc = lib().orc_objective_batch(C.c_int(ntime), C.c_int(ncell), _p(P, C.c_double), _p(E, C.c_double),
                                   _p(area, C.c_double), _p(ndays, C.c_int), _p(region, C.c_int), C.c_int(nreg),
                                   _p(params, C.c_double), C.c_int(nsets), _p(qobs, C.c_double),
                                   C.c_int(warmup or 0), C.c_int(flags), C.c_int(nthreads),
                                   _p(crit, C.c_double), _p(of, C.c_double), _p(qt, C.c_double))
    if rc != 0:
        raise ValueError(f"orc_objective_batch failed: {rc}")
    return dict(crit=crit, of=of, qt=qt)


def d8_plan(flowdir):
    """Topology plan of a TauDEM D8 grid (int16 codes 1..8, anything else nodata)."""
    d = np.ascontiguousarray(flowdir, dtype=np.int16)
    nrow, ncol = d.shape
    n = nrow * ncol
    receiver, indeg, level, order = (np.empty(n, np.int32) for _ in range(4))
    level_off = np.empty(n + 2, np.int32)
    inmask, contam = np.empty(n, np.uint8), np.empty(n, np.uint8)
    nord = C.c_int32(0)
    nlev = lib().orc_d8_plan(C.c_int(nrow), C.c_int(ncol), _p(d, C.c_int16), _p(receiver, C.c_int32),
                             _p(inmask, C.c_uint8), _p(indeg, C.c_int32), _p(level, C.c_int32),
                             _p(order, C.c_int32), _p(level_off, C.c_int32), _p(contam, C.c_uint8),
                             C.byref(nord))
    if nlev < 0:
        raise ValueError(f"orc_d8_plan failed: {nlev}")
    return dict(receiver=receiver, inmask=inmask, indeg=indeg, level=level, order=order[:nord.value].copy(),
                level_off=level_off[:nlev + 1].copy(), contam=contam, nlevels=nlev)


def aread8(flowdir, W, f32=False, contcheck=True):
    d = np.ascontiguousarray(flowdir, dtype=np.int16)
    nrow, ncol = d.shape
    W = _f64(W).reshape(-1)
    A = np.empty(nrow * ncol)
    lib().orc_aread8(C.c_int(nrow), C.c_int(ncol), _p(d, C.c_int16), _p(W, C.c_double), C.c_int(int(f32)),
                     C.c_int(int(contcheck)), _p(A, C.c_double))
    return A.reshape(nrow, ncol)


def route(flowdir, QS, src_cell, poly_off, poly_cells, f32=False, contcheck=True, nthreads=1):
    """Routing_GR2MSemiDistr month loop. QS: [nsub, ntime]. Returns QR [nsub, ntime] unrounded."""
    d = np.ascontiguousarray(flowdir, dtype=np.int16)
    nrow, ncol = d.shape
    QS = _f64(QS)
    nsub, ntime = QS.shape
    src, poff, pc = _i32(src_cell), _i32(poly_off), _i32(poly_cells)
    QR = np.empty((nsub, ntime))
    rc = lib().orc_route(C.c_int(nrow), C.c_int(ncol), _p(d, C.c_int16), C.c_int(ntime), C.c_int(nsub),
                         _p(QS, C.c_double), _p(src, C.c_int32), _p(poff, C.c_int32), _p(pc, C.c_int32),
                         C.c_int(int(f32)), C.c_int(int(contcheck)), C.c_int(nthreads), _p(QR, C.c_double))
    if rc != 0:
        raise ValueError(f"orc_route failed: {rc}")
    return QR
