"""Measured parity margins of the CUDA path against the CPU oracle (GPU only; writes one JSON line per case).
Error measure = max |gpu - oracle| / max(|oracle|, 1), the bar of the tests is 1e-10."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
import oracle
from gr2msemidistr_b200 import capi, synth

g.build(); capi.set_device(0)
rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0), initial=0.0))

def sim_case(name, ncell, ntime, nreg, seed, flags=0):
    P, E, area, nd = synth.forcing(ncell, ntime, seed=seed)
    region = (np.arange(ncell) * 7 % nreg).astype(np.int32)
    par = synth.param_sets(3, nreg, seed=seed + 1)[2 if nreg > 1 else 0]
    ref = oracle.run(P, E, area, nd, region, nreg, par, flags=flags & 1, nthreads=8)
    with capi.Basin(P, E, area, nd, region, nreg) as b:
        got = b.run(par, flags=flags)
    out = {"case": name, "PR_bit_exact": bool(np.array_equal(got["PR"], ref["PR"]))}
    out.update({k: rel(got[k], ref[k]) for k in ("AE", "SM", "RU", "QS", "S_end", "R_end")})
    print(json.dumps(out), flush=True)

def obj_case(name, ncell, ntime, nsets, warmup, seed, flags=0):
    P, E, area, nd = synth.forcing(ncell, ntime, seed=seed)
    region = np.zeros(ncell, np.int32)
    par = synth.param_sets(nsets, seed=seed + 1)
    base = oracle.run(P, E, area, nd, region, 1, par[0])
    qobs = synth.qobs_from(oracle.outlet_optim(base["QS"]))
    ref = oracle.objective_batch(P, E, area, nd, region, 1, par, qobs, warmup, flags=(flags & 1) | oracle.SINK_UNROUNDED, nthreads=8)
    with capi.Basin(P, E, area, nd, region, 1) as b:
        got = b.objective_batch(par, qobs, warmup, capi.OF_KGE, flags | capi.SINK_UNROUNDED, want_qt=True)
        rnd = b.objective_batch(par, qobs, warmup, capi.OF_KGE, flags, want_qt=True)
    refr = oracle.objective_batch(P, E, area, nd, region, 1, par, qobs, warmup, flags=flags & 1, nthreads=8)
    ok = np.isfinite(ref["crit"])
    print(json.dumps({"case": name, "outlet_series": rel(got["qt"], ref["qt"]), "KGE_NSE_RMSE": rel(got["crit"][ok], ref["crit"][ok]),
                      "rounded_outlet_values_differing": float(np.mean(rnd["qt"] != refr["qt"])),
                      "rounded_objective_values_differing": float(np.mean(rnd["of"] != refr["of"][:, 0]))}), flush=True)

sim_case("C1 shape 13 x 432, set0 (fast step)", 13, 432, 1, 0)
sim_case("C1 shape 13 x 432, set0 (literal step)", 13, 432, 1, 0, capi.MATH_LITERAL)
sim_case("C1 shape, airGR single-precision exponent", 13, 432, 1, 0, capi.PERC_F32_EXPONENT)
sim_case("C3 shape 3600 x 480, 14 regions", 3600, 480, 14, 3)
obj_case("C2 shape 13 x 264 x 10000 sets, WarmUp 36", 13, 264, 10000, 36, 0)
obj_case("C2 shape, literal step, 2000 sets", 13, 264, 2000, 36, 0, capi.MATH_LITERAL)
obj_case("C4 slice 2000 x 480 x 64 sets", 2000, 480, 64, 36, 5)
