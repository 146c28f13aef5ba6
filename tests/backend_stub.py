"""Oracle-backed stand-in for gr2msemidistr_b200.capi, used ONLY by CPU tests of the host-side
wrappers (date subsetting, parameter tables, rounding, list assembly, SCE-UA, sharding).  It has the
same Basin / Plan surface as capi but computes with the CPU oracle -- test infrastructure, never a
product path: the wrappers always call gr2msemidistr_b200.capi; tests swap it with pytest's monkeypatch."""
import numpy as np

import oracle

OF_NAMES = {"KGE": 0, "NSE": 1, "RMSE": 2}
SINK_UNROUNDED = 8


class Basin:
    def __init__(self, P, E, area, ndays, region_id, nreg):
        self.P, self.E = np.asarray(P, float), np.asarray(E, float)
        self.area, self.ndays, self.region, self.nreg = area, ndays, region_id, nreg

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def close(self):
        pass

    def run(self, params, S0=None, R0=None, flags=0, outputs=None):
        return oracle.run(self.P, self.E, self.area, self.ndays, self.region, self.nreg, params, S0, R0, flags & 1)

    def objective_batch(self, params, qobs, warmup=0, which=1, flags=0, want_crit=True, want_qt=False):
        r = oracle.objective_batch(self.P, self.E, self.area, self.ndays, self.region, self.nreg, params, qobs,
                                   warmup, flags & 9, nthreads=4)
        return dict(of=r["of"][:, which].copy(), crit=r["crit"], qt=r["qt"])


def flowdir_from_dem(dem, valid, want_filled=False):
    from oracle import d8prep
    z = np.where(valid, np.asarray(dem, float), np.nan)
    filled, fd = d8prep.fill_and_flowdir(z)
    return (fd, filled) if want_filled else fd


class Plan:
    def __init__(self, flowdir, flags=0):
        self.d = np.asarray(flowdir, np.int16)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def route(self, QS, src, off, cells, flags=0):
        return oracle.route(self.d, QS, src, off, cells, f32=bool(flags & 1), contcheck=not (flags & 2), nthreads=4)
