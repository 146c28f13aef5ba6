"""The R binding's C glue (r_package/src/gr2m_shim.c), compiled and EXECUTED without R.

R is not in the image, so tests/mock_R/ provides a minimal stand-in for R's C API (column-major matrices,
NA_INTEGER, named lists, external pointers with finalizers, the registered .Call table) and shim_driver.c calls
the shim through that table exactly as `.Call(C_gr2m_run, ...)` in r_package/R/gr2m_b200.R does.  What is checked:

  not gpu   the shim compiles warning-free against the API it uses, every routine the R code .Call()s is
            registered with the argument count the R code passes, NAMESPACE exports what the reference's does;
  gpu       the driver runs on the B200 and its outputs equal the ctypes path's bit for bit: the column-major /
            row-major flips of C_wfa_route and C_wfa_flowdir, NA handling, the IniState hand-off, the batched
            objective with one candidate per column, the finalizer.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from gr2msemidistr_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MOCK = os.path.join(HERE, "mock_R")
SHIM = os.path.join(ROOT, "r_package", "src", "gr2m_shim.c")
CFLAGS = ["-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type"]   # DL_FUNC casts are R's own idiom


def build_driver(outdir):
    import __graft_entry__ as g
    g.build()
    libdir = os.path.join(ROOT, "gr2msemidistr_b200")
    exe = os.path.join(outdir, "shim_driver")
    objs = []
    for src in (SHIM, os.path.join(MOCK, "mock_R.c"), os.path.join(MOCK, "shim_driver.c")):
        o = os.path.join(outdir, os.path.basename(src) + ".o")
        subprocess.check_call(["gcc", *CFLAGS, "-I", MOCK, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", o])
        objs.append(o)
    subprocess.check_call(["gcc", "-o", exe, *objs, "-L", libdir, "-lgr2m_b200", f"-Wl,-rpath,{libdir}", "-lm"])
    return exe


def test_shim_compiles_and_links_against_the_mock_api(tmp_path):
    exe = build_driver(str(tmp_path))
    assert os.path.exists(exe)


def test_registered_routines_match_the_r_code_and_namespace():
    shim = open(SHIM).read()
    rcode = open(os.path.join(ROOT, "r_package", "R", "gr2m_b200.R")).read()
    table = dict(re.findall(r'\{"(C_\w+)", \(DL_FUNC\)&\1, (\d+)\}', shim))
    assert set(table) == {"C_gr2m_run", "C_gr2m_basin", "C_gr2m_objective", "C_wfa_route", "C_wfa_flowdir", "C_wfa_router",
                          "C_gr2m_objective_routed"}
    # every .Call in the R code names a registered routine and passes as many arguments as registered
    calls = re.findall(r"\.Call\((C_\w+)((?:[^()]|\([^()]*(?:\([^()]*\)[^()]*)*\))*)\)", rcode)
    assert calls
    for name, args in calls:
        depth, n = 0, 0
        for ch in args:
            depth += ch in "([{"
            depth -= ch in ")]}"
            n += ch == "," and depth == 0
        assert name in table and int(table[name]) == n, (name, n, table.get(name))
    # the reference's NAMESPACE:3-6 exports four functions; the drop-in package exports them all
    ns = open(os.path.join(ROOT, "r_package", "NAMESPACE")).read()
    for fn in ("Create_Forcing_Inputs", "Run_GR2MSemiDistr", "Optim_GR2MSemiDistr", "Routing_GR2MSemiDistr"):
        assert f"export({fn})" in ns and re.search(rf"^{fn} <- function\(", rcode, re.M), fn
    # exported signatures keep the reference's argument names (R/Run_GR2MSemiDistr.R:48-56, R/Optim_GR2MSemiDistr.R:41-51,
    # R/Routing_GR2MSemiDistr.R:35-41, R/Create_Forcing_Inputs.R:42-54)
    sigs = {"Run_GR2MSemiDistr": ["Data", "Subbasins", "RunIni", "RunEnd", "WarmUp", "Parameters", "IniState", "Save", "Update"],
            "Optim_GR2MSemiDistr": ["Data", "Subbasins", "RunIni", "RunEnd", "WarmUp", "Parameters", "Parameters.Min", "Parameters.Max",
                                    "Max.Functions", "Optimization", "No.Optim"],
            "Routing_GR2MSemiDistr": ["Model", "Subbasins", "Dem", "AcumIni", "AcumEnd", "Save", "Update"],
            "Create_Forcing_Inputs": ["Subbasins", "Precip", "PotEvap", "Qobs", "DateIni", "DateEnd", "IniGrids", "Save", "Update",
                                      "Resolution", "Buffer", "Members"]}
    for fn, want in sigs.items():
        m = re.search(rf"^{fn} <- function\(([^{{]*)\)\s*\{{", rcode, re.M | re.S)
        got = [a.split("=")[0].strip() for a in re.sub(r"\s+", " ", m.group(1)).split(",")]
        assert got[:len(want)] == want, (fn, got)


@pytest.mark.gpu
def test_shim_outputs_equal_the_ctypes_path(tmp_path):
    from gr2msemidistr_b200 import capi
    assert capi.device_count() >= 1
    capi.set_device(0)
    exe = build_driver(str(tmp_path))
    d = str(tmp_path)
    ntime, ncell, nreg, nsets, warm, which, flags = 50, 45, 2, 19, 6, 0, 0
    P, E, area, nd = synth.forcing(ncell, ntime, seed=81)
    region = (np.arange(ncell) % nreg).astype(np.int32)
    par = synth.param_sets(nsets, nreg, seed=82)
    qobs = synth.qobs_from(np.full(ntime, 80.0))
    # a small DEM with a nodata corner and a pit, as raster::as.matrix would give it: [nrow, ncol] column-major, NA = nodata
    rng = np.random.default_rng(83)
    nrow, ncol = 37, 53
    rr, cc = np.meshgrid(np.arange(nrow), np.arange(ncol), indexing="ij")
    dem = 500.0 - 3.0 * rr - 2.0 * cc + rng.uniform(0, 4, (nrow, ncol))
    dem[10:13, 20:24] -= 30.0
    valid = np.ones((nrow, ncol), bool)
    valid[:6, :9] = False
    src, off, cells = synth.block_polygons(nrow, ncol, 3, 4, margin=1)
    rt_ntime, rt_nsub = 7, len(src)
    QS = rng.uniform(0, 300, (rt_nsub, rt_ntime))

    def put(name, a, dt):
        np.ascontiguousarray(a, dtype=dt).tofile(os.path.join(d, name))
    # a second geometry with one polygon per simulated subbasin, for the routed objective
    n2r, n2c = 48, 60
    d8b = synth.d8_grid(n2r, n2c, seed=84).numpy()
    src2, off2, cells2 = synth.block_polygons(n2r, n2c, 5, 9, margin=1)
    assert len(src2) == ncell
    gauge = 31
    put("dims.bin", [ntime, ncell, nreg, nsets, warm, which, nrow, ncol, rt_ntime, rt_nsub, len(cells), flags,
                     n2r, n2c, len(cells2), gauge], np.int32)
    put("rt2_flowdir.bin", np.where(d8b < 1, np.iinfo(np.int32).min, d8b.astype(np.int32)).T, np.int32)   # column-major, NA = INT_MIN
    put("rt2_src.bin", src2, np.int32); put("rt2_off.bin", off2, np.int32); put("rt2_cells.bin", cells2, np.int32)
    put("P.bin", P, np.float64); put("E.bin", E, np.float64)                 # [ncell, ntime] C order == [ntime x ncell] column-major
    put("area.bin", area, np.float64); put("ndays.bin", nd, np.int32); put("region.bin", region, np.int32)
    put("par0.bin", par[0], np.float64)
    put("params.bin", par, np.float64)                                         # [nsets, 4 nreg] C order == one candidate per column
    put("qobs.bin", qobs, np.float64)
    put("dem.bin", np.where(valid, dem, np.nan).T, np.float64)                 # column-major [nrow x ncol]
    put("rt_QS.bin", QS, np.float64); put("rt_src.bin", src, np.int32); put("rt_off.bin", off, np.int32)
    put("rt_cells.bin", cells if len(cells) else [0], np.int32)
    res = subprocess.run([exe, d], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "shim driver ok" in res.stdout

    def get(name, dt, shape):
        return np.fromfile(os.path.join(d, name), dtype=dt).reshape(shape)
    ref = capi.run_once(P, E, area, nd, region, nreg, par[0])
    for k in ("PR", "AE", "SM", "RU", "QS"):
        assert np.array_equal(get(f"out_{k}.bin", np.float64, (ncell, ntime)), ref[k]), k
    for k in ("S_end", "R_end"):
        assert np.array_equal(get(f"out_{k}.bin", np.float64, ncell), ref[k]), k
    again = capi.run_once(P, E, area, nd, region, nreg, par[0], S0=ref["S_end"], R0=ref["R_end"])
    assert np.array_equal(get("out_QS_restart.bin", np.float64, (ncell, ntime)), again["QS"])
    with capi.Basin(P, E, area, nd, region, nreg) as b:
        of = b.objective_batch(par, qobs, warm, which, flags)["of"]
    assert np.array_equal(get("out_of.bin", np.float64, nsets), of, equal_nan=True)
    with capi.Plan(d8b) as plan2, plan2.router(src2, off2, cells2) as router2, capi.Basin(P, E, area, nd, region, nreg) as b:
        ofr = b.objective_routed_batch(router2, par, qobs, gauge, warm, which, flags)["of"]
    assert np.array_equal(get("out_of_routed.bin", np.float64, nsets), ofr, equal_nan=True)
    fd = capi.flowdir_from_dem(np.where(valid, dem, 0.0), valid)
    got_fd = get("out_flowdir.bin", np.int32, (ncol, nrow)).T                  # column-major integer matrix, NA = INT_MIN
    assert np.array_equal(np.where(got_fd == np.iinfo(np.int32).min, -32768, got_fd), fd.astype(np.int32))
    with capi.Plan(fd) as plan:
        QR = plan.route(QS, src, off, cells)
    assert np.array_equal(get("out_QR.bin", np.float64, (rt_nsub, rt_ntime)), QR, equal_nan=True)
