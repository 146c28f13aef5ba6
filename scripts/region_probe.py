"""Calibration sweep with several regions (subbasins grouped by region): table-mode passes vs the general kernel.
    python scripts/region_probe.py [ncell nsets nreg]      # GR2M_CALIB_GENERIC=1 forces the general kernel"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gr2msemidistr_b200 import capi, synth
ncell, nsets, nreg = (int(x) for x in (sys.argv[1:4] + ["20000", "1250", "3"][len(sys.argv) - 1:]))
capi.set_device(0)
P, E, area, nd = synth.forcing(ncell, 480)
region = np.minimum(np.arange(ncell) * nreg // ncell, nreg - 1).astype(np.int32)
par = synth.param_sets(nsets, nreg)
qobs = synth.qobs_from(np.full(480, 500.0))
with capi.Basin(P, E, area, nd, region, nreg) as b:
    ms = []
    for _ in range(4):
        got = b.objective_batch(par, qobs, 36, capi.OF_KGE, 0, want_crit=False)
        ms.append(b.last_kernel_ms())
    print("%d subbasins x %d sets x 480 months, %d regions (%s): kernel %.2f ms = %.4g cell-months/s, checksum %.6f" % (
        ncell, nsets, nreg, "general kernel" if os.environ.get("GR2M_CALIB_GENERIC") else "table-mode passes",
        min(ms[1:]), ncell * nsets * 480.0 / (min(ms[1:]) * 1e-3), float(np.nansum(got["of"]))))
