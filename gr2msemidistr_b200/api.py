"""Host-side mirror of the reference's exported R functions, over the C ABI.

Same names, argument meaning, return structures and rounding as the reference:

    Run_GR2MSemiDistr      R/Run_GR2MSemiDistr.R:48
    Optim_GR2MSemiDistr    R/Optim_GR2MSemiDistr.R:41
    Routing_GR2MSemiDistr  R/Routing_GR2MSemiDistr.R:35
    Create_Forcing_Inputs  R/Create_Forcing_Inputs.R:42   (host-only, not on the hot path)

R is not available in this image, so this Python layer stands where the R wrappers of R/ (shipped in
this repo for a machine that has R) stand: it does the date subsetting, the
parameter table, the output rounding/assembly -- and nothing numerical.  All arithmetic of the
hot path happens in libgr2m_b200.so; if the library or a GPU is missing these functions raise.

Conventions: `Data` is a pandas DataFrame with the columns of Create_Forcing_Inputs
(DatesR, P_<comid>..., E_<comid>..., optional Q) (Create:185-189); `Subbasins` is a gis.Subbasins
(Area, Region, COMID, polygons); matrices in the returned dicts are [ntime, nsub] like the R ones.
"""
from __future__ import annotations

import datetime as _dt
import os
import time

import numpy as np

from . import capi as _capi
from .rround import rround

_DATE_FORMATS = ("%Y-%m-%d", "%Y/%m/%d", "%d-%m-%Y", "%d/%m/%Y")   # tryFormats of Run:74


def _parse_dates(col):
    out = []
    for v in col:
        if isinstance(v, (_dt.datetime, _dt.date)):
            out.append(_dt.date(v.year, v.month, v.day))
            continue
        if hasattr(v, "to_pydatetime"):
            d = v.to_pydatetime()
            out.append(_dt.date(d.year, d.month, d.day))
            continue
        s = str(v).split(" ")[0]
        for f in _DATE_FORMATS:
            try:
                out.append(_dt.datetime.strptime(s, f).date())
                break
            except ValueError:
                continue
        else:
            raise ValueError(f"character string is not in a standard unambiguous format: {v!r}")
    return out


def _days_in_month(d: _dt.date) -> int:
    """numberOfDays(), Run:108-117."""
    nxt = _dt.date(d.year + (d.month == 12), d.month % 12 + 1, 1)
    return (nxt - _dt.date(d.year, d.month, 1)).days


def _subset(Data, RunIni, RunEnd):
    """Ind_run / Database of Run:74-79: rows from the month RunIni to the month RunEnd ('mm/yyyy')."""
    dates = _parse_dates(Data["DatesR"])
    key = [d.strftime("%m/%Y") for d in dates]
    if RunIni not in key or RunEnd not in key:
        raise ValueError(f"RunIni/RunEnd ({RunIni}, {RunEnd}) not found in Data$DatesR")
    i0, i1 = key.index(RunIni), key.index(RunEnd)
    if i1 < i0:
        raise ValueError("RunEnd precedes RunIni")
    return dates[i0:i1 + 1], slice(i0, i1 + 1)


def _forcing_matrices(Data, rows, nsub):
    """Column i+1 = P of subbasin i, column i+1+nsub = E (Run:99).  Returns [nsub, ntime] arrays."""
    cols = list(Data.columns)
    if len(cols) < 1 + 2 * nsub:
        raise ValueError("Data must hold DatesR, nsub precipitation and nsub evapotranspiration columns")
    arr = Data.iloc[rows, 1:1 + 2 * nsub].to_numpy(dtype=np.float64)
    return np.ascontiguousarray(arr[:, :nsub].T), np.ascontiguousarray(arr[:, nsub:].T)


def _regions(Subbasins):
    zone = sorted(set(Subbasins.Region.tolist()))          # sort(unique(region)), Run:82
    rid = np.array([zone.index(r) for r in Subbasins.Region.tolist()], np.int32)
    return zone, rid


def _drop_warmup(x, WarmUp):
    """R's x[-WarmUp:-1] / x[-WarmUp:-1, ]: drop the first WarmUp rows (Run:240-251)."""
    return x if WarmUp is None else x[int(WarmUp):]


def Run_GR2MSemiDistr(Data, Subbasins, RunIni, RunEnd, WarmUp=None, Parameters=None, IniState=None,
                      Save=False, Update=False, flags=0):
    """GR2M for n subbasins.  Returns dict(SINK?, QS, RU, PR, AE, SM, Dates, COMID, EndState)."""
    be = _capi
    t0 = time.perf_counter()
    area = np.asarray(Subbasins.Area, np.float64)
    comid = np.asarray(Subbasins.COMID)
    nsub = len(area)
    dates, rows = _subset(Data, RunIni, RunEnd)
    ntime = len(dates)
    zone, rid = _regions(Subbasins)
    nreg = len(zone)
    Parameters = np.asarray(Parameters, np.float64)
    if Parameters.size != 4 * nreg:
        raise ValueError(f"Parameters must hold 4 values per region ({4 * nreg}): X1.., X2.., Fpp.., Fpe..")
    P, E = _forcing_matrices(Data, rows, nsub)
    ndays = np.array([_days_in_month(d) for d in dates], np.int32)
    S0 = R0 = None
    if IniState is not None:                                  # Run:161-171 reads Store$Prod / Store$Rout
        S0 = np.array([s["Store"]["Prod"] for s in IniState], np.float64)
        R0 = np.array([s["Store"]["Rout"] for s in IniState], np.float64)
    print(f"Running GR2M model for {nsub} subbasins")          # message(), Run:122
    with be.Basin(P, E, area, ndays, rid, nreg) as b:
        o = b.run(Parameters, S0, R0, flags)
    # airGR-shaped StateEnd so that EndState round-trips as IniState (Run:224, 161-171)
    EndState = [{"Store": {"Prod": float(o["S_end"][i]), "Rout": float(o["R_end"][i]), "Exp": None},
                 "UH": {"UH1": None, "UH2": None}} for i in range(nsub)]
    if nsub == 1:                                             # Run:193-199
        pr, ae, sm, ru = (rround(o[k].T, 1) for k in ("PR", "AE", "SM", "RU"))
        qs = rround((area[0] * ru) / (86.4 * ndays[:, None]), 1)
        qt = qs[:, 0]
    else:                                                     # Run:231-236
        pr, ae, sm, ru, qs = (rround(o[k].T, 1) for k in ("PR", "AE", "SM", "RU", "QS"))
        qt = rround(np.sum(qs.astype(np.longdouble), axis=1).astype(np.float64), 2)
    qo = Data["Q"].to_numpy(dtype=np.float64)[rows] if "Q" in Data.columns else None
    pr, ae, sm, ru, qs, qt = (_drop_warmup(x, WarmUp) for x in (pr, ae, sm, ru, qs, qt))
    dts = _drop_warmup(dates, WarmUp)
    if qo is not None:
        qo = _drop_warmup(qo, WarmUp)
    Ans = {}
    if qo is not None:                                        # Run:254-266
        Ans["SINK"] = {"sim": rround(qt, 3), "obs": rround(qo, 3)}
    Ans.update(QS=qs, RU=ru, PR=pr, AE=ae, SM=sm, Dates=dts, COMID=comid, EndState=EndState)
    if Save:
        _save_run(Ans, RunEnd, Update)
    print("Done!")
    Ans["elapsed_s"] = time.perf_counter() - t0                # tic()/toc(), Run:65,342
    return Ans


def _write_table(path, mat, dates, names):
    """write.table(x, sep='\\t') of a data frame with Date row names (Run:336-339)."""
    with open(path, "w") as f:
        f.write("\t".join(f'"{n}"' for n in names) + "\n")
        for d, row in zip(dates, mat):
            f.write(f'"{d.isoformat()}"\t' + "\t".join(repr(float(v)).rstrip("0").rstrip(".") if np.isfinite(v) else "NA"
                                                       for v in row) + "\n")


def _read_table(path):
    with open(path) as f:
        names = [s.strip('"') for s in f.readline().rstrip("\n").split("\t")]
        dates, rows = [], []
        for line in f:
            parts = line.rstrip("\n").split("\t")
            dates.append(_dt.date.fromisoformat(parts[0].strip('"')))
            rows.append([np.nan if p == "NA" else float(p) for p in parts[1:]])
    return names, dates, np.array(rows)


def _month_shift(today: _dt.date, k: int) -> str:
    y, m = today.year, today.month - k
    while m < 1:
        y, m = y - 1, m + 12
    return f"{y:04d}{m:02d}"


def _save_run(Ans, RunEnd, Update, outdir="./Outputs", today=None):
    """Run:280-340.  Update mode appends to last month's files and removes them."""
    os.makedirs(outdir, exist_ok=True)
    today = today or _dt.date.today()
    for key in ("PR", "AE", "SM", "RU"):
        names = [f"{key}_{c}" for c in Ans["COMID"]]
        if Update:
            old = os.path.join(outdir, f"{key}_GR2MSemiDistr_{_month_shift(today, 2)}.txt")
            new = os.path.join(outdir, f"{key}_GR2MSemiDistr_{_month_shift(today, 1)}.txt")
            _, d_old, m_old = _read_table(old)
            _write_table(new, np.vstack([m_old, Ans[key]]), d_old + list(Ans["Dates"]), names)
            os.remove(old)
        else:
            mnyr = _dt.datetime.strptime("01/" + RunEnd, "%d/%m/%Y").strftime("%Y%m")
            _write_table(os.path.join(outdir, f"{key}_GR2MSemiDistr_{mnyr}.txt"), Ans[key], Ans["Dates"], names)


class _Objective:
    """OFUN of Optim:125-217 with the forcing resident on the GPU; evaluates many sets per call."""

    def __init__(self, be, P, E, area, ndays, rid, nreg, qobs, WarmUp, Optimization, fixed_idx, fixed_val, npar,
                 flags=0, routed=None):
        self.b = be.Basin(P, E, area, ndays, rid, nreg)
        self.routed = routed               # (plan, router, gauge index): score the routed discharge of one polygon
        self.qobs, self.warmup = qobs, int(WarmUp or 0)
        self.which = _capi.OF_NAMES[Optimization]
        self.fixed_idx, self.fixed_val, self.npar, self.flags = fixed_idx, fixed_val, npar, flags
        self.free_idx = np.setdiff1d(np.arange(npar), fixed_idx)
        self.calls = 0

    def full(self, X):
        """Splice the non-optimised regions' parameters back (Optim:128-134)."""
        X = np.atleast_2d(np.asarray(X, np.float64))
        full = np.empty((X.shape[0], self.npar))
        full[:, self.free_idx] = X
        full[:, self.fixed_idx] = self.fixed_val
        return full

    def __call__(self, X):
        X = np.atleast_2d(np.asarray(X, np.float64))
        self.calls += X.shape[0]
        if self.routed is not None:
            return self.b.objective_routed_batch(self.routed[1], self.full(X), self.qobs, self.routed[2], self.warmup,
                                                 self.which, self.flags, want_crit=False)["of"]
        return self.b.objective_batch(self.full(X), self.qobs, self.warmup, self.which, self.flags,
                                      want_crit=False)["of"]

    def close(self):
        self.b.close()
        if self.routed is not None:
            self.routed[1].close()
            self.routed[0].close()


def Optim_GR2MSemiDistr(Data, Subbasins, RunIni, RunEnd, WarmUp=None, Parameters=None, Parameters_Min=None,
                        Parameters_Max=None, Max_Functions=5000, Optimization="NSE", No_Optim=None,
                        ngs=5, restarts=8, seed=0, flags=0, shard=None, Gauge=None, Dem=None, FlowDir=None):
    """SCE-UA calibration.  Returns dict(Param, Value) rounded as Optim:232-244.

    `Gauge` (a COMID) switches the objective from the outlet sum of all subbasins (the reference's OFUN) to the ROUTED
    discharge of that subbasin's polygon: every candidate is simulated, routed over `Dem` (or the given `FlowDir`) by the
    month-invariant operator and scored there -- Run -> Routing -> hydroGOF in one device call per batch.

    `Max_Functions` and `ngs` mean what they mean to rtop::sceua in the reference (evaluation budget and number
    of complexes of ONE search; ngs = 5 is rtop's default).  The GPU batch is filled by `restarts` independent
    searches advancing in lock-step, the best of which is returned (restart 0 starts from `Parameters`); the
    objective is evaluated for all of their candidate points at once (see sceua.py);
    `shard=(rank, world, allgather)` splits each batch across ranks."""
    from .sceua import sceua_batched
    be = _capi
    if Optimization not in _capi.OF_NAMES:
        raise ValueError("Optimization must be 'NSE', 'KGE' or 'RMSE'")
    t0 = time.perf_counter()
    area = np.asarray(Subbasins.Area, np.float64)
    nsub = len(area)
    dates, rows = _subset(Data, RunIni, RunEnd)
    zone, rid = _regions(Subbasins)
    nreg = len(zone)
    Parameters = np.asarray(Parameters, np.float64)
    if Parameters.size != 4 * nreg:
        raise ValueError(f"Parameters must hold 4 values per region ({4 * nreg})")
    if "Q" not in Data.columns:
        raise ValueError("Data$Q (observed discharge) is required for calibration")
    P, E = _forcing_matrices(Data, rows, nsub)
    ndays = np.array([_days_in_month(d) for d in dates], np.int32)
    qobs = Data["Q"].to_numpy(dtype=np.float64)[rows]
    # optimised vs fixed regions (Optim:78-90): v.param = rep(sort(unique(region)), 4)
    vparam = np.array(zone * 4)
    fixed = np.isin(vparam, list(No_Optim)) if No_Optim is not None else np.zeros(4 * nreg, bool)
    n_opt_reg = nreg - len(set(No_Optim) & set(zone)) if No_Optim is not None else nreg
    lower = np.repeat(np.asarray(Parameters_Min, np.float64), n_opt_reg)   # rep(Parameters.Min, each=...)
    upper = np.repeat(np.asarray(Parameters_Max, np.float64), n_opt_reg)
    x0 = Parameters[~fixed]
    if lower.size != x0.size or upper.size != x0.size:
        raise ValueError("Parameters.Min / Parameters.Max must hold 4 values (X1, X2, fp, fpe)")
    routed = None
    if Gauge is not None:
        if Dem is None:
            raise ValueError("Gauge needs Dem (and optionally FlowDir) to route the simulated discharge")
        if FlowDir is None:
            z = np.asarray(Dem.data[0], np.float64)
            valid = np.isfinite(z) & (z > -1e30)
            if Dem.nodata is not None:
                valid &= ~np.isclose(z, Dem.nodata, rtol=1e-6)
            FlowDir = be.flowdir_from_dem(z, valid)
        src, off, cells = routing_geometry(Subbasins, Dem)
        gauge = int(np.flatnonzero(np.asarray(Subbasins.COMID) == Gauge)[0])
        plan = be.Plan(FlowDir)
        routed = (plan, plan.router(src, off, cells), gauge)
    obj = _Objective(be, P, E, area, ndays, rid, nreg, qobs, WarmUp, Optimization, np.flatnonzero(fixed),
                     Parameters[fixed], 4 * nreg, flags, routed)
    print(f"Optimizing {Optimization} with SCE-UA")             # Optim:223
    try:
        res = sceua_batched(obj, x0, lower, upper, maxn=Max_Functions, ngs=ngs, restarts=restarts, seed=seed, shard=shard)
    finally:
        obj.close()
    fo = rround(res["value"], 3) if Optimization == "RMSE" else rround(1.0 - res["value"], 3)   # Optim:232-236
    par = Parameters.copy()
    par[~fixed] = rround(res["par"], 3)                         # Optim:237-243
    Ans = {"Param": par, "Value": float(fo), "evaluations": res["evaluations"], "generations": res["generations"],
           "elapsed_s": time.perf_counter() - t0}
    names = np.repeat(["X1=", "X2=", "fpr=", "fpe="], nreg)
    print("Optimization results:")
    print("======================")
    print([f"{n}{v}" for n, v in zip(names, Ans["Param"])])
    print(f"{Optimization}={Ans['Value']}")
    print("Done!")
    return Ans


def routing_geometry(Subbasins, Dem):
    """Month-invariant inputs of the routing seam (Routing:77-81,119-121): the DEM cell of each polygon's
    point on surface and the cells with coverage fraction 1.  Dem is a gis.Raster."""
    from . import gis
    pts = [gis.point_on_surface(r) for r in Subbasins.polygons]
    src = gis.cell_of_points(Dem, [p[0] for p in pts], [p[1] for p in pts])
    cells = [gis.interior_cells(Dem, r) for r in Subbasins.polygons]
    off = np.concatenate([[0], np.cumsum([len(c) for c in cells])]).astype(np.int32)
    flat = np.concatenate(cells).astype(np.int32) if off[-1] else np.empty(0, np.int32)
    return src, off, flat


def Routing_GR2MSemiDistr(Model, Subbasins, Dem, AcumIni=None, AcumEnd=None, Save=False, Update=False,
                          FlowDir=None, flags=0):
    """Weighted-Flow-Accumulation routing of Model$QS.  Returns dict(QR, Dates, COMID).

    `FlowDir` (int16 TauDEM codes) replaces the pitremove + D8Flowdir calls (Routing:88-93) when given;
    otherwise wfa_flowdir_from_dem derives one from `Dem` on the device (own flat rule, documented)."""
    be = _capi
    t0 = time.perf_counter()
    QS = np.asarray(Model["QS"], np.float64)
    dates = list(Model["Dates"])
    if QS.ndim == 1:
        QS = QS[None, :]
    if not ((AcumIni is None and AcumEnd is None) or QS.shape[0] == 1):          # Routing:57-68
        key = [d.strftime("%d/%m/%Y") for d in dates]
        i0, i1 = key.index("01/" + AcumIni), key.index("01/" + AcumEnd)
        QS, dates = QS[i0:i1 + 1], dates[i0:i1 + 1]
    ntime, nsub = QS.shape
    if FlowDir is None:
        z = np.asarray(Dem.data[0], np.float64)
        valid = np.isfinite(z) & (z > -1e30)
        if Dem.nodata is not None:
            valid &= ~np.isclose(z, Dem.nodata, rtol=1e-6)
        FlowDir = be.flowdir_from_dem(z, valid)
    src, off, cells = routing_geometry(Subbasins, Dem)
    if np.any(src < 0):
        raise ValueError("a subbasin's point on surface falls outside the DEM")
    print(f"Routing streamflows for {nsub} sub-basins")        # Routing:103
    with be.Plan(FlowDir) as plan:
        QR = plan.route(np.ascontiguousarray(QS.T), src, off, cells, flags)
    qr = rround(QR.T, 1)                                        # Routing:126-132
    Ans = {"QR": qr, "Dates": dates, "COMID": np.asarray(Subbasins.COMID)}
    if Save:
        _save_routing(Ans, Update)
    print("Done!")
    Ans["elapsed_s"] = time.perf_counter() - t0
    return Ans


def _save_routing(Ans, Update, outdir="./Outputs", today=None):
    """Routing:140-160."""
    os.makedirs(outdir, exist_ok=True)
    today = today or _dt.date.today()
    names = [f"QR_{c}" for c in Ans["COMID"]]
    if Update:
        old = os.path.join(outdir, f"QR_GR2MSemiDistr_{_month_shift(today, 2)}.txt")
        new = os.path.join(outdir, f"QR_GR2MSemiDistr_{_month_shift(today, 1)}.txt")
        _, d_old, m_old = _read_table(old)
        _write_table(new, np.vstack([m_old, Ans["QR"]]), d_old + list(Ans["Dates"]), names)
        os.remove(old)
    else:
        last = Ans["Dates"][-1]
        _write_table(os.path.join(outdir, f"QR_GR2MSemiDistr_{last.strftime('%Y%m')}.txt"), Ans["QR"], Ans["Dates"], names)


def Create_Forcing_Inputs(Subbasins, Precip, PotEvap, Qobs=None, DateIni=None, DateEnd=None, Resolution=0.01,
                          start=(1981, 1)):
    """Areal-mean forcing table in airGR layout (Create:42-218): DatesR, P_<comid>.., E_<comid>.., Q.

    Precip / PotEvap are gis.Raster bricks with one band per month starting at `start`.  The zonal mean
    is taken on the `Resolution` grid the reference resamples to (nearest neighbour) by sampling cell
    centres -- an approximation of exactextractr's exact coverage fractions.  Host only."""
    import pandas as pd

    from . import gis
    nb = Precip.data.shape[0]
    dates = [_dt.date(start[0] + (start[1] - 1 + i) // 12, (start[1] - 1 + i) % 12 + 1, 1) for i in range(nb)]
    key = [d.strftime("%Y/%m/%d") for d in dates]
    i0 = key.index(DateIni) if DateIni else 0
    i1 = key.index(DateEnd) if DateEnd else nb - 1
    mp = rround(gis.zonal_means(Precip, Subbasins, Resolution), 1)[:, i0:i1 + 1]     # Create:114
    me = rround(gis.zonal_means(PotEvap, Subbasins, Resolution), 1)[:, i0:i1 + 1]
    cols = {"DatesR": dates[i0:i1 + 1]}
    for i, c in enumerate(Subbasins.COMID):
        cols[f"P_{c}"] = mp[i]
    for i, c in enumerate(Subbasins.COMID):
        cols[f"E_{c}"] = me[i]
    if Qobs is not None:
        cols["Q"] = np.asarray(Qobs, np.float64)[i0:i1 + 1]
    return pd.DataFrame(cols)
