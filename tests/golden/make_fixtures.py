"""Build the committed C1/C2 fixtures from the reference's bundled inputs (run in the dev container,
where /root/reference exists; the GPU box only sees the .npz files written here).

    python tests/golden/make_fixtures.py

c1_inputs.npz   decoded raw/roi.*, raw/dem.tif, raw/pisco_pr.tif, raw/pisco_pe.tif, raw/qobs.txt:
                the Create_Forcing_Inputs table (zonal means, approximation documented in gis.py),
                subbasin attributes, DEM, the D8 grid derived from it (d8prep), the routing geometry.
c1_golden.npz   CPU-oracle outputs on those inputs for the documented parameter set
                c(10.976, 0.665, 1.186, 1.169) (R/Run_GR2MSemiDistr.R:33) and for a small calibration
                batch inside the documented bounds (R/Optim_GR2MSemiDistr.R:29-31).
The reference ships inputs only (no expected outputs), so these golden vectors pin the ORACLE's
answers, not R's: parity with the reference itself stays unpinned.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
RAW = "/root/reference/raw"

import oracle  # noqa: E402
from gr2msemidistr_b200 import api, gis, synth  # noqa: E402
from oracle import d8prep  # noqa: E402


def main():
    sb = gis.read_shapefile(os.path.join(RAW, "roi"))
    dem = gis.read_tiff(os.path.join(RAW, "dem.tif"))
    pr = gis.read_tiff(os.path.join(RAW, "pisco_pr.tif"))
    pe = gis.read_tiff(os.path.join(RAW, "pisco_pe.tif"))
    q = []
    for line in open(os.path.join(RAW, "qobs.txt")):
        parts = line.split()
        if len(parts) >= 2:
            q.append(np.nan if parts[1] == "NA" else float(parts[1]))
    qobs = np.array(q)
    assert len(qobs) == 432 and np.isnan(qobs).sum() == 12
    data = api.Create_Forcing_Inputs(sb, pr, pe, qobs, start=(1981, 1))
    nsub = len(sb)
    P = data.iloc[:, 1:1 + nsub].to_numpy().T.copy()
    E = data.iloc[:, 1 + nsub:1 + 2 * nsub].to_numpy().T.copy()
    demv = dem.data[0]
    dem_i16 = np.where(demv > -1e30, np.rint(demv), -32768).astype(np.int16)
    assert np.array_equal(np.where(demv > -1e30, demv, -32768), dem_i16)       # DEM is integer valued
    _, fdir = d8prep.fill_and_flowdir(demv, dem.nodata)
    src, off, cells = api.routing_geometry(sb, dem)
    nd = synth.ndays(432)
    region = np.zeros(nsub, np.int32)
    np.savez_compressed(os.path.join(HERE, "c1_inputs.npz"), P=P, E=E, area=sb.Area, comid=sb.COMID, qobs=qobs,
                        ndays=nd, dem=dem_i16, dem_geo=np.array([dem.x0, dem.y0, dem.dx, dem.dy]), flowdir=fdir,
                        src_cell=src, poly_off=off, poly_cells=cells,
                        poly_xy=np.concatenate([np.column_stack([np.full(len(r[0]), i), r[0]]) for i, r in enumerate(sb.polygons)]))
    par0 = np.array(synth.SET0)
    sim = oracle.run(P, E, sb.Area, nd, region, 1, par0)
    qt_run = oracle.outlet_run(sim["QS"], sim["RU"], sb.Area, nd)
    qs_r = oracle.r_round(sim["QS"], 1)
    QR = oracle.route(fdir, qs_r, src, off, cells)
    QR32 = oracle.route(fdir, qs_r, src, off, cells, f32=True)
    pl = oracle.d8_plan(fdir)
    par = synth.param_sets(64)
    i0, i1 = 0, 264                                            # 01/1981..12/2002, WarmUp 36 (Optim:26-32)
    ob = oracle.objective_batch(P[:, i0:i1], E[:, i0:i1], sb.Area, nd[i0:i1], region, 1, par, qobs[i0:i1], warmup=36)
    ob_u = oracle.objective_batch(P[:, i0:i1], E[:, i0:i1], sb.Area, nd[i0:i1], region, 1, par, qobs[i0:i1], warmup=36,
                                  flags=oracle.SINK_UNROUNDED)
    np.savez_compressed(os.path.join(HERE, "c1_golden.npz"), params0=par0, PR=sim["PR"], AE=sim["AE"], SM=sim["SM"],
                        RU=sim["RU"], QS=sim["QS"], S_end=sim["S_end"], R_end=sim["R_end"], qt_run=qt_run, QR=QR,
                        QR_f32=QR32, level=pl["level"], order=pl["order"], level_off=pl["level_off"],
                        contam=pl["contam"], params=par, of=ob["of"], crit=ob["crit"], qt=ob["qt"],
                        crit_unrounded=ob_u["crit"], qt_unrounded=ob_u["qt"])
    write_pin_inputs(P, E, sb, qobs, fdir, src, qs_r)
    for f in ("c1_inputs.npz", "c1_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    print("levels", pl["nlevels"], "src", src, "interior cells per polygon", np.diff(off))
    print("P mean/max", P.mean(), P.max(), "E mean", E.mean())
    print("outlet sim mean", qt_run.mean(), "obs mean", np.nanmean(qobs), "NSE/KGE set0", ob["crit"][0])
    print("QR finite share", np.isfinite(QR).mean(), "QR max", np.nanmax(np.where(np.isfinite(QR), QR, np.nan)))


def pin_param_sets():
    """Parameter sets the R dump script evaluates: the documented set, the corners of the calibration box
    (Optim:29-31), sets that trigger the P/X1 > 13 clamp and exact decimal ties, and a few random ones."""
    fixed = [synth.SET0, [1000.0, 1.0, 1.0, 1.0], [1.0, 0.01, 0.8, 0.8], [2000.0, 2.0, 1.2, 1.2], [1.0, 2.0, 1.2, 0.8],
             [2000.0, 0.01, 0.8, 1.2], [25.0, 0.9, 1.15, 0.85], [350.0, 0.7, 1.05, 1.1]]
    return np.vstack([np.array(fixed), np.round(synth.param_sets(8, seed=77), 3)])


def write_pin_inputs(P, E, sb, qobs, fdir, src, qs_r):
    """tests/golden/pin_inputs/*: what scripts/dump_reference_goldens.R feeds to R / airGR / hydroGOF / TauDEM."""
    import gzip
    d = os.path.join(HERE, "pin_inputs")
    os.makedirs(d, exist_ok=True)
    nsub, ntime = P.shape
    with open(os.path.join(d, "data.csv"), "w") as f:
        f.write(",".join(["DatesR"] + [f"P_{c}" for c in sb.COMID] + [f"E_{c}" for c in sb.COMID] + ["Q"]) + "\n")
        for t in range(ntime):
            row = [f"{1981 + t // 12:04d}-{t % 12 + 1:02d}-01"] + [repr(float(v)) for v in P[:, t]] + [repr(float(v)) for v in E[:, t]]
            row.append("NA" if np.isnan(qobs[t]) else repr(float(qobs[t])))
            f.write(",".join(row) + "\n")
    with open(os.path.join(d, "subbasins.csv"), "w") as f:
        f.write("COMID,Area,Region\n")
        for c, a, r in zip(sb.COMID, sb.Area, sb.Region):
            f.write(f"{int(c)},{float(a)!r},{r}\n")
    with open(os.path.join(d, "param_sets.csv"), "w") as f:
        f.write("X1,X2,Fpp,Fpe\n")
        for p in pin_param_sets():
            f.write(",".join(repr(float(v)) for v in p) + "\n")
    with open(os.path.join(d, "config.csv"), "w") as f:
        f.write("key,value\nn_sim_sets,3\nrun_ini,01/1981\nrun_end,12/2016\nrun_warmup,12\noptim_end,12/2002\noptim_warmup,36\n")
    rng = np.random.default_rng(5)
    k = rng.integers(0, 400000, 400).astype(np.float64)
    x = np.concatenate([[0.0, 0.05, 0.15, 0.25, 0.35, 29.65, 2.5, 0.125, 0.375, 1e-9, 3839.95, -0.15, -7.25, 1234.5675, 0.0045],
                        (k + 0.5) / 10.0, 1.186 * (rng.integers(0, 40000, 300) / 10.0), 1.169 * (rng.integers(0, 3000, 300) / 10.0),
                        rng.uniform(0, 4000, 300), (k[:200] + 0.5) / 100.0, (k[:200] + 0.5) / 1000.0])
    with open(os.path.join(d, "round_cases.csv"), "w") as f:
        f.write("x\n" + "\n".join(repr(float(v)) for v in x) + "\n")
    with gzip.GzipFile(os.path.join(d, "flowdir.csv.gz"), "wb", mtime=0) as f:
        f.write("\n".join(",".join(str(int(v)) for v in row) for row in fdir).encode() + b"\n")
    months = [0, 2, 431]
    with open(os.path.join(d, "route_weights.csv"), "w") as f:
        f.write("cell," + ",".join(f"m{m}" for m in months) + "\n")
        for i in range(nsub):
            f.write(f"{int(src[i])}," + ",".join(repr(float(qs_r[i, m])) for m in months) + "\n")


if __name__ == "__main__":
    main()
