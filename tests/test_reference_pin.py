"""Pin of the oracle -- and of the CUDA path -- to the REAL reference (R + airGR + hydroGOF [+ GR2MSemiDistr, TauDEM]).

The reference ships no expected outputs and none of its arithmetic can run in the build image, so the
oracle is a restatement ("parity unpinned").  `scripts/dump_reference_goldens.R` produces
`tests/golden/ref/*.csv` on any machine with R; once those files are committed the tests below compare

  * the CPU oracle (always), and
  * the CUDA path through the C ABI (`-m gpu`)

against them at the north-star tolerance (1e-10 relative FP64; rounded outputs and R's round() bit-exact).
While the files are absent every reference-consuming test SKIPS with a message naming the script -- that is
what "unpinned" looks like in the test report.  The consumer code itself is exercised on every run against
a reference-shaped directory written from the oracle (`test_pin_consumer_self_check`), so the day real files
arrive the comparison cannot fail for a parsing reason.

One decision is waiting for the pin: airGR's percolation exponent.  The library's default is the exact cube
root; flag GR2M_PERC_F32_EXPONENT gives `(double)(1.f/3.f)`, what gfortran makes of a literal `(1./3.)`.  The two
differ by <= 7e-9 relative -- more than the 1e-10 bar -- so exactly one of them can pass `check_series`; the test
says which, and the default must be flipped if it is the flag.
"""
import csv
import gzip
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PIN = os.path.join(HERE, "golden", "pin_inputs")
REF = os.path.join(HERE, "golden", "ref")
TOL = 1e-10
UNPINNED = ("parity unpinned: tests/golden/ref/ does not exist -- run `Rscript scripts/dump_reference_goldens.R` on a "
            "machine with R >= 4.0, airGR and hydroGOF (optionally GR2MSemiDistr and TauDEM) and commit its output")


# ---- readers -----------------------------------------------------------------------------------------
def _num(v):
    return np.nan if v in ("NA", "NaN", "") else float(v)


def read_table(path):
    """CSV with a header row -> (names, float array [rows, cols]); non-numeric columns come back as NaN."""
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        rows = list(csv.reader(f))
    names = rows[0]

    def conv(v):
        try:
            return _num(v.strip())
        except ValueError:
            return np.nan
    return names, np.array([[conv(v) for v in r] for r in rows[1:]], np.float64).reshape(len(rows) - 1, len(names))


def load_inputs():
    names, d = read_table(os.path.join(PIN, "data.csv"))
    nsub = (len(names) - 2) // 2
    with open(os.path.join(PIN, "subbasins.csv")) as f:
        sub = list(csv.DictReader(f))
    area = np.array([float(r["Area"]) for r in sub])
    zones = sorted({r["Region"] for r in sub})
    region = np.array([zones.index(r["Region"]) for r in sub], np.int32)
    sets = read_table(os.path.join(PIN, "param_sets.csv"))[1]
    with open(os.path.join(PIN, "config.csv")) as f:
        cfg = {r["key"]: r["value"] for r in csv.DictReader(f)}
    with open(os.path.join(PIN, "data.csv")) as f:
        dates = [r.split(",")[0] for r in f.read().splitlines()[1:]]
    from gr2msemidistr_b200 import synth
    nd = synth.ndays(len(dates))                      # table starts 1981-01 (checked below)
    assert dates[0] == "1981-01-01"
    P = np.ascontiguousarray(d[:, 1:1 + nsub].T)
    E = np.ascontiguousarray(d[:, 1 + nsub:1 + 2 * nsub].T)
    ncal = dates.index(f"{cfg['optim_end'][3:]}-{cfg['optim_end'][:2]}-01") + 1
    return dict(P=P, E=E, qobs=d[:, -1].copy(), area=area, region=region, nreg=len(zones), sets=sets, ndays=nd,
                nsim=int(cfg["n_sim_sets"]), ncal=ncal, warm=int(cfg["optim_warmup"]), run_warm=int(cfg["run_warmup"]),
                round_x=read_table(os.path.join(PIN, "round_cases.csv"))[1][:, 0])


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0), initial=0.0))


# ---- comparisons (shared by the oracle and the CUDA leg) ------------------------------------------------
def check_series(ref_dir, inp, run, f32_flag):
    """run(P, E, area, ndays, region, nreg, par, flags) -> dict of [nsub, ntime] series + S_end/R_end.  Returns the
    worst relative error of the default and of the single-precision-exponent variant."""
    worst = {0: 0.0, f32_flag: 0.0}
    ndays_ref = read_table(os.path.join(ref_dir, "ref_ndays.csv"))[1][:, 0]
    assert np.array_equal(ndays_ref, inp["ndays"]), "days per month differ from Run_GR2MSemiDistr.R:108-117"
    ends = read_table(os.path.join(ref_dir, "ref_airgr_endstate.csv"))[1]
    for k in range(inp["nsim"]):
        ref = {v: read_table(os.path.join(ref_dir, f"ref_airgr_set{k:02d}_{v}.csv"))[1].T for v in ("PR", "AE", "SM", "RU", "QS")}
        e = ends[ends[:, 0] == k]
        for fl in worst:
            got = run(inp["P"], inp["E"], inp["area"], inp["ndays"], inp["region"], inp["nreg"], inp["sets"][k], fl)
            assert np.array_equal(got["PR"], ref["PR"]), f"set {k}: forcing correction round(f*x, 1) is not bit-exact vs R"
            err = max(relerr(got[v], ref[v]) for v in ("AE", "SM", "RU", "QS"))
            err = max(err, relerr(got["S_end"], e[:, 2]), relerr(got["R_end"], e[:, 3]))
            worst[fl] = max(worst[fl], err)
    return worst[0], worst[f32_flag]


def assert_exponent_verdict(e_default, e_f32):
    assert min(e_default, e_f32) <= TOL, (
        f"neither percolation exponent reproduces airGR: exact 1/3 -> {e_default:.3g}, single-precision literal -> {e_f32:.3g}")
    assert e_default <= TOL, (
        f"airGR behaves like the single-precision exponent (error {e_f32:.3g} with GR2M_PERC_F32_EXPONENT, {e_default:.3g} "
        "without): make the flag the default in include/gr2m_b200.h / oracle and regenerate tests/golden")


def check_objective(ref_dir, inp, objective, flags=0):
    """objective(P, E, area, ndays, region, nreg, sets, qobs, warmup, flags) -> dict(crit [n,3], of [n,3], qt [n, ntime])."""
    n = inp["ncal"]
    got = objective(inp["P"][:, :n], inp["E"][:, :n], inp["area"], inp["ndays"][:n], inp["region"], inp["nreg"], inp["sets"],
                    inp["qobs"][:n], inp["warm"], flags)
    ref = read_table(os.path.join(ref_dir, "ref_objective.csv"))[1]
    crit, of = ref[:, 1:4], ref[:, 4:7]
    assert np.array_equal(np.isnan(got["crit"]), np.isnan(crit)), "criteria are NA for different sets than in hydroGOF"
    ok = ~np.isnan(crit)
    assert relerr(got["crit"][ok], crit[ok]) <= 1e-8          # criteria of sinks rounded to 0.01: a tie flips 1e-2/mean
    d = np.abs(got["of"] - of)[ok]
    assert np.all(d <= 1e-3 + 1e-9) and np.mean(d > 1e-9) <= 0.05
    for k in range(inp["nsim"]):
        sink = read_table(os.path.join(ref_dir, f"ref_sink_set{k:02d}.csv"))[1][:, 0]
        dq = np.abs(got["qt"][k] - sink)
        assert np.all(dq <= 0.01 + 1e-9) and np.mean(dq > 1e-9) <= 0.01


def check_round(ref_dir, inp, r_round):
    ref = read_table(os.path.join(ref_dir, "ref_round.csv"))[1]
    assert np.array_equal(ref[:, 0], inp["round_x"])
    for dig in (1, 2, 3):
        assert np.array_equal(r_round(ref[:, 0], dig), ref[:, dig]), f"R round(x, {dig}) differs"


def check_gof_degenerate(ref_dir, gof):
    with open(os.path.join(ref_dir, "ref_gof_degenerate.csv")) as f:
        ref = {r["case"]: [_num(r[k]) for k in ("KGE", "NSE", "RMSE")] for r in csv.DictReader(f)}
    nan = np.nan
    obs5 = np.array([3.0, 5, 4, 9, 1])
    cases = {"regular": ([1.0, 2, 3, 4, 6], [1.0, 3, 2, 5, 5]), "constant_sim": (np.full(5, 2.0), obs5),
             "constant_obs": (obs5, np.full(5, 2.0)), "zero_obs": (obs5, np.zeros(5)), "one_pair": ([1.0, 2, 3], [nan, 2.5, nan]),
             "two_pairs": ([1.0, 2, 3], [nan, 2.5, 2.0]), "no_pair": ([1.0, 2], [nan, nan]), "identical": (obs5, obs5)}
    for name, (sim, obs) in cases.items():
        crit = gof(np.asarray(sim, float), np.asarray(obs, float))
        want = np.array(ref[name])
        assert np.array_equal(np.isnan(crit), np.isnan(want)), name
        ok = ~np.isnan(want)
        assert relerr(crit[ok], want[ok]) <= 1e-12, name


def check_package_run(ref_dir, inp, run_wrapper):
    """Tier B: Run_GR2MSemiDistr's own (rounded) return values for set 0."""
    if not os.path.exists(os.path.join(ref_dir, "ref_pkg_run_QS.csv")):
        pytest.skip("tier B of the pin absent (GR2MSemiDistr was not installed where the goldens were dumped)")
    got = run_wrapper(inp)
    for k in ("QS", "RU", "PR", "AE", "SM"):
        ref = read_table(os.path.join(ref_dir, f"ref_pkg_run_{k}.csv"))[1]
        d = np.abs(got[k] - ref)
        assert d.shape == ref.shape and np.all(d <= 0.1 + 1e-9) and np.mean(d > 1e-9) <= 1e-3, k
    sink = read_table(os.path.join(ref_dir, "ref_pkg_run_SINK.csv"))[1]
    assert np.all(np.abs(got["SINK"]["sim"] - sink[:, 0]) <= 0.1 * inp["P"].shape[0] + 1e-9)
    assert np.array_equal(np.isnan(got["SINK"]["obs"]), np.isnan(sink[:, 1]))


def check_aread8(ref_dir, aread8_f32):
    """Tier C: TauDEM's float32 accumulation grids, bit for bit (nodata / contaminated cells as NaN)."""
    files = sorted(f for f in os.listdir(ref_dir) if f.startswith("ref_ad8_month"))
    if not files:
        pytest.skip("tier C of the pin absent (TauDEM was not available where the goldens were dumped)")
    with gzip.open(os.path.join(PIN, "flowdir.csv.gz"), "rt") as f:
        fd = np.array([[int(v) for v in line.split(",")] for line in f.read().splitlines()], np.int16)
    _, w = read_table(os.path.join(PIN, "route_weights.csv"))
    for j, fn in enumerate(files):
        with gzip.open(os.path.join(ref_dir, fn), "rt") as f:
            ref = np.array([[_num(v) for v in line.split(",")] for line in f.read().splitlines()])
        W = np.zeros(fd.size)
        W[w[:, 0].astype(int)] = w[:, 1 + j]
        got = aread8_f32(fd, W.reshape(fd.shape))
        got = np.where(got < 0, np.nan, got)                        # AreaD8 nodata (-1) reads back as NA
        assert np.array_equal(np.isnan(got), np.isnan(ref)), fn
        assert np.array_equal(got[~np.isnan(ref)].astype(np.float32), ref[~np.isnan(ref)].astype(np.float32)), fn


# ---- adapters --------------------------------------------------------------------------------------------
def oracle_adapters(orc):
    def run(P, E, area, nd, region, nreg, par, fl):
        return orc.run(P, E, area, nd, region, nreg, par, flags=fl)

    def objective(P, E, area, nd, region, nreg, sets, qobs, warm, flags):
        return orc.objective_batch(P, E, area, nd, region, nreg, sets, qobs, warm, flags=flags, nthreads=4)
    return run, objective


def gpu_adapters(capi):
    def run(P, E, area, nd, region, nreg, par, fl):
        return capi.run_once(P, E, area, nd, region, nreg, par, flags=fl)

    def objective(P, E, area, nd, region, nreg, sets, qobs, warm, flags):
        with capi.Basin(P, E, area, nd, region, nreg) as b:
            res = {w: b.objective_batch(sets, qobs, warm, w, flags, want_qt=True) for w in (0, 1, 2)}
        return dict(crit=res[0]["crit"], of=np.column_stack([res[w]["of"] for w in (0, 1, 2)]), qt=res[0]["qt"])
    return run, objective


def write_reference_shaped_dir(path, inp, orc):
    """What dump_reference_goldens.R writes (tier A), produced from the oracle: exercises the consumer code only."""
    os.makedirs(path, exist_ok=True)

    def wr(name, header, rows):
        with open(os.path.join(path, name), "w") as f:
            f.write(",".join(header) + "\n")
            for r in rows:
                f.write(",".join("NA" if isinstance(v, float) and np.isnan(v) else (v if isinstance(v, str) else f"{v:.17g}") for v in r) + "\n")
    nsub, ntime = inp["P"].shape
    ends = []
    for k in range(inp["nsim"]):
        sim = orc.run(inp["P"], inp["E"], inp["area"], inp["ndays"], inp["region"], inp["nreg"], inp["sets"][k])
        for v in ("PR", "AE", "SM", "RU", "QS"):
            wr(f"ref_airgr_set{k:02d}_{v}.csv", [f"V{i + 1}" for i in range(nsub)], sim[v].T.tolist())
        ends += [[float(k), float(i), float(sim["S_end"][i]), float(sim["R_end"][i])] for i in range(nsub)]
    wr("ref_airgr_endstate.csv", ["set", "subbasin", "Prod", "Rout"], ends)
    wr("ref_ndays.csv", ["ndays"], [[float(v)] for v in inp["ndays"]])
    n = inp["ncal"]
    ob = orc.objective_batch(inp["P"][:, :n], inp["E"][:, :n], inp["area"], inp["ndays"][:n], inp["region"], inp["nreg"],
                             inp["sets"], inp["qobs"][:n], inp["warm"])
    wr("ref_objective.csv", ["set", "KGE", "NSE", "RMSE", "of_KGE", "of_NSE", "of_RMSE"],
       [[float(k)] + ob["crit"][k].tolist() + ob["of"][k].tolist() for k in range(len(inp["sets"]))])
    for k in range(inp["nsim"]):
        wr(f"ref_sink_set{k:02d}.csv", ["sink"], [[float(v)] for v in ob["qt"][k]])
    x = inp["round_x"]
    wr("ref_round.csv", ["x", "r1", "r2", "r3"], np.column_stack([x] + [orc.r_round(x, d) for d in (1, 2, 3)]).tolist())
    nan = np.nan
    obs5 = np.array([3.0, 5, 4, 9, 1])
    cases = {"regular": ([1.0, 2, 3, 4, 6], [1.0, 3, 2, 5, 5]), "constant_sim": (np.full(5, 2.0), obs5),
             "constant_obs": (obs5, np.full(5, 2.0)), "zero_obs": (obs5, np.zeros(5)), "one_pair": ([1.0, 2, 3], [nan, 2.5, nan]),
             "two_pairs": ([1.0, 2, 3], [nan, 2.5, 2.0]), "no_pair": ([1.0, 2], [nan, nan]), "identical": (obs5, obs5)}
    wr("ref_gof_degenerate.csv", ["case", "KGE", "NSE", "RMSE"],
       [[name] + [float(v) for v in orc.gof(np.asarray(s, float), np.asarray(o, float))[1]] for name, (s, o) in cases.items()])


# ---- tests -----------------------------------------------------------------------------------------------
def test_pin_inputs_are_the_committed_fixture():
    inp = load_inputs()
    c1 = np.load(os.path.join(HERE, "golden", "c1_inputs.npz"))
    assert np.array_equal(inp["P"], c1["P"]) and np.array_equal(inp["E"], c1["E"])
    assert np.array_equal(inp["qobs"], c1["qobs"], equal_nan=True) and np.array_equal(inp["area"], c1["area"])
    assert np.array_equal(inp["ndays"], c1["ndays"]) and inp["ncal"] == 264 and inp["warm"] == 36
    assert np.array_equal(inp["sets"][0], [10.976, 0.665, 1.186, 1.169])          # Run_GR2MSemiDistr.R:33
    script = open(os.path.join(os.path.dirname(HERE), "scripts", "dump_reference_goldens.R")).read()
    for name in ("data.csv", "subbasins.csv", "param_sets.csv", "config.csv", "round_cases.csv", "flowdir.csv.gz", "route_weights.csv"):
        assert os.path.exists(os.path.join(PIN, name)) and name in script
    for out in ("ref_airgr_set%02d_", "ref_airgr_endstate.csv", "ref_ndays.csv", "ref_objective.csv", "ref_sink_set%02d.csv",
                "ref_gof_degenerate.csv", "ref_round.csv", "ref_pkg_run_", "ref_ad8_month%02d.csv.gz"):
        assert out in script, f"the dump script no longer writes {out}"


def test_pin_consumer_self_check(oracle, tmp_path):
    """The comparison code run against a reference-shaped directory written from the oracle itself: proves the
    readers, layouts and adapters work; says nothing about the reference (that is what tests/golden/ref is for)."""
    inp = load_inputs()
    write_reference_shaped_dir(str(tmp_path), inp, oracle)
    run, objective = oracle_adapters(oracle)
    e0, e1 = check_series(str(tmp_path), inp, run, oracle.PERC_F32_EXPONENT)
    assert e0 == 0.0 and 1e-10 < e1 < 1e-7            # the two exponents are told apart by the 1e-10 bar
    assert_exponent_verdict(e0, e1)
    check_objective(str(tmp_path), inp, objective)
    check_round(str(tmp_path), inp, oracle.r_round)
    check_gof_degenerate(str(tmp_path), lambda s, o: oracle.gof(s, o)[1])


def test_oracle_against_reference_goldens(oracle):
    if not os.path.isdir(REF):
        pytest.skip(UNPINNED)
    inp = load_inputs()
    run, objective = oracle_adapters(oracle)
    assert_exponent_verdict(*check_series(REF, inp, run, oracle.PERC_F32_EXPONENT))
    check_objective(REF, inp, objective)
    check_round(REF, inp, oracle.r_round)
    check_gof_degenerate(REF, lambda s, o: oracle.gof(s, o)[1])


def test_oracle_aread8_against_taudem(oracle):
    if not os.path.isdir(REF):
        pytest.skip(UNPINNED)
    check_aread8(REF, lambda fd, W: np.where(np.isnan(a := oracle.aread8(fd, W, f32=True)), -1.0, a))


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    assert m.device_count() >= 1, "gpu tests need a CUDA device"
    m.set_device(0)
    return m


@pytest.mark.gpu
def test_gpu_consumer_self_check(capi, oracle, tmp_path):
    """CUDA path through the pin's own comparison code against an oracle-written directory (always runs)."""
    inp = load_inputs()
    write_reference_shaped_dir(str(tmp_path), inp, oracle)
    run, objective = gpu_adapters(capi)
    e0, e1 = check_series(str(tmp_path), inp, run, capi.PERC_F32_EXPONENT)
    assert e0 <= TOL and 1e-10 < e1 < 1e-7
    check_objective(str(tmp_path), inp, objective)
    check_round(str(tmp_path), inp, lambda x, d: capi.debug_math(5, x, np.full(x.size, float(d))))


@pytest.mark.gpu
def test_gpu_against_reference_goldens(capi):
    if not os.path.isdir(REF):
        pytest.skip(UNPINNED)
    inp = load_inputs()
    run, objective = gpu_adapters(capi)
    assert_exponent_verdict(*check_series(REF, inp, run, capi.PERC_F32_EXPONENT))
    check_objective(REF, inp, objective)
    check_round(REF, inp, lambda x, d: capi.debug_math(5, x, np.full(x.size, float(d))))


@pytest.mark.gpu
def test_gpu_run_wrapper_against_the_reference_package(capi):
    if not os.path.isdir(REF):
        pytest.skip(UNPINNED)
    import pandas as pd

    from gr2msemidistr_b200 import api, gis

    def run_wrapper(inp):
        data = pd.read_csv(os.path.join(PIN, "data.csv"))
        with open(os.path.join(PIN, "subbasins.csv")) as f:
            sub = list(csv.DictReader(f))
        sb = gis.Subbasins(inp["area"], np.array([r["Region"] for r in sub]), np.array([int(r["COMID"]) for r in sub]), [])
        return api.Run_GR2MSemiDistr(data, sb, "01/1981", "12/2016", WarmUp=inp["run_warm"], Parameters=inp["sets"][0])
    check_package_run(REF, load_inputs(), run_wrapper)
