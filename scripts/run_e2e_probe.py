"""gr2m_run end to end on 1e5 x 480 with fresh pageable result arrays (what Run_GR2MSemiDistr's .Call sees):
    python scripts/run_e2e_probe.py            # GR2M_NO_THP=1 / GR2M_NO_STAGING=1 for the A/B"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gr2msemidistr_b200 import capi, synth
capi.set_device(0)
ncell, nt = 100000, 480
P, E, area, nd = synth.forcing(ncell, nt)
par = np.array(synth.SET0)
with capi.Basin(P, E, area, nd, np.zeros(ncell, np.int32), 1) as b:
    b.run(par)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        out = b.run(par)
        ts.append(time.perf_counter() - t0)
        del out
    print("gr2m_run %d x %d: best %.1f ms, median %.1f ms = %.3g cell-months/s (THP advice %s, thp=%s)" % (
        ncell, nt, min(ts) * 1e3, sorted(ts)[2] * 1e3, ncell * nt / sorted(ts)[2],
        "off" if os.environ.get("GR2M_NO_THP") else "on", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()))
