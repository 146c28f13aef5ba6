"""Memory-safety stand-in for compute-sanitizer (refused on the GPU pool).

GR2M_GUARD=1 makes the library allocate every handle buffer between two guard bands (checked when the buffer is
released) and pre-fill its payload with 0xFF bytes.  Each kernel family is run once in that mode, in a fresh process
(the switch is read when the library loads), on shapes that are ragged against every tile / stage / warp size in the
kernels; the run must leave all guard bands intact AND still match the CPU oracle, which it cannot if a kernel reads
a word it never wrote (the 0xFF fill turns those into NaN / -1)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, sys
import numpy as np
sys.path.insert(0, %(root)r)
import oracle
from gr2msemidistr_b200 import capi, synth
capi.set_device(0)
TOL = 1e-10
def rel(a, b): return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0), initial=0.0))
rng = np.random.default_rng(5)
report = {}
# recursion: simulation kernel, table-mode and general calibration kernels, objective -- ragged sizes
for ncell, ntime, nreg, nsets in ((77, 61, 1, 19), (45, 50, 3, 9), (1, 13, 1, 3), (333, 75, 1, 61), (2100, 27, 1, 10)):
    P, E, area, nd = synth.forcing(ncell, ntime, seed=ncell)
    region = rng.integers(0, nreg, ncell).astype(np.int32)
    par = synth.param_sets(nsets, nreg, seed=ntime)
    ref = oracle.run(P, E, area, nd, region, nreg, par[1])
    qobs = synth.qobs_from(oracle.outlet_optim(ref["QS"]))
    want = oracle.objective_batch(P, E, area, nd, region, nreg, par, qobs, 3, flags=oracle.SINK_UNROUNDED, nthreads=4)
    with capi.Basin(P, E, area, nd, region, nreg) as b:
        got = b.run(par[1])
        assert np.array_equal(got["PR"], ref["PR"])
        for k in ("AE", "SM", "RU", "QS", "S_end", "R_end"):
            assert rel(got[k], ref[k]) <= TOL, (k, ncell)
        for fl in (0, capi.MATH_LITERAL):
            ob = b.objective_batch(par, qobs, 3, capi.OF_KGE, fl | capi.SINK_UNROUNDED, want_qt=True)
            assert rel(ob["qt"], want["qt"]) <= TOL, ("qt", ncell, fl)
            ok = np.isfinite(want["crit"])
            assert np.array_equal(np.isfinite(ob["crit"]), ok) and rel(ob["crit"][ok], want["crit"][ok]) <= TOL
report["recursion"] = capi.guard_counts()
# routing: plan, dense sweep (FP64 and float32), operator, device pit fill -- a grid that is ragged against 32 x 32 tiles
for n0, n1 in ((61, 83), (130, 97)):
    d8 = synth.d8_grid(n0, n1, seed=n0).numpy()
    src, off, cells = synth.block_polygons(n0, n1, 3, 4, margin=1)
    QS = rng.uniform(0, 300, (len(src), 9))
    with capi.Plan(d8) as plan:
        pl, ref_pl = plan.export(), oracle.d8_plan(d8)
        for k in ("receiver", "inmask", "indeg", "level", "order", "level_off", "contam"):
            assert np.array_equal(pl[k], ref_pl[k]), k
        for fl in (0, capi.WFA_ROUTE_DENSE, capi.WFA_F32 | capi.WFA_ROUTE_DENSE, capi.WFA_LEVEL_SYNC | capi.WFA_ROUTE_DENSE,
                   capi.WFA_BY_BASIN | capi.WFA_ROUTE_DENSE, capi.WFA_DATAFLOW | capi.WFA_ROUTE_DENSE):
            QR = plan.route(QS, src, off, cells, fl)
            QR_ref = oracle.route(d8, QS, src, off, cells, f32=bool(fl & capi.WFA_F32))
            fin = np.isfinite(QR_ref)
            assert np.array_equal(np.isfinite(QR), fin), fl
            assert np.array_equal(QR[fin], QR_ref[fin]), fl
    z = 300.0 - 2.0 * np.arange(n0)[:, None] - 1.5 * np.arange(n1)[None, :] + rng.uniform(0, 6, (n0, n1))
    z[20:25, 30:37] -= 40.0
    valid = np.ones((n0, n1), bool); valid[:5, :7] = False
    from oracle import d8prep
    _, fd_ref = d8prep.fill_and_flowdir(np.where(valid, z, -9999.0), -9999.0)
    assert np.array_equal(capi.flowdir_from_dem(z, valid), fd_ref)
report["routing"] = capi.guard_counts()
print("GUARD_REPORT " + json.dumps(report))
'''


def test_every_kernel_family_under_guard_bands():
    import __graft_entry__ as g
    g.build()
    env = dict(os.environ, GR2M_GUARD="1")
    res = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "[gr2m guard]" not in res.stderr, res.stderr[-2000:]
    line = [x for x in res.stdout.splitlines() if x.startswith("GUARD_REPORT ")][-1]
    rep = json.loads(line[len("GUARD_REPORT "):])
    checked, bad = rep["routing"]
    assert bad == 0 and rep["recursion"][1] == 0
    assert rep["recursion"][0] > 50 and checked > 200          # the mode was really on: hundreds of buffers went through the check


def test_guard_mode_catches_an_overrun():
    """The checker itself: a deliberate one-element overrun (debug hook) must be reported."""
    import __graft_entry__ as g
    g.build()
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from gr2msemidistr_b200 import capi\n"
            "capi.set_device(0)\n"
            "import ctypes as C\n"
            "rc = capi.lib().gr2m_debug_overrun(C.c_int(1000), C.c_int(1))\n"
            "print('COUNTS', rc, *capi.guard_counts())\n") % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, GR2M_GUARD="1"), timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    counts = [x for x in res.stdout.splitlines() if x.startswith("COUNTS")][-1].split()
    assert counts[1] == "0" and int(counts[2]) >= 1 and int(counts[3]) == 1
    assert "[gr2m guard]" in res.stderr
    res = subprocess.run([sys.executable, "-c", code.replace("C.c_int(1))", "C.c_int(0))")], capture_output=True, text=True,
                         env=dict(os.environ, GR2M_GUARD="1"), timeout=300)
    counts = [x for x in res.stdout.splitlines() if x.startswith("COUNTS")][-1].split()
    assert int(counts[3]) == 0
