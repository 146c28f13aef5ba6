"""Time the batched WFA sweep variants on a synthetic D8 grid (development aid, GPU only)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
from gr2msemidistr_b200 import capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8000)
ap.add_argument("--batches", default="16,32")
ap.add_argument("--flags", default="0,128")
ap.add_argument("--terrain", default="tilted", choices=["tilted", "filled"])
ap.add_argument("--reps", type=int, default=3, help="sweeps per variant (1 under ncu)")
a = ap.parse_args()
g.build(); capi.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
if a.terrain == "filled":
    z, valid = synth.fractal_dem(a.n, a.n, device=dev)
    v8 = valid.to(torch.uint8).contiguous(); del valid
    d8 = torch.empty((a.n, a.n), dtype=torch.int16, device=dev)
    capi.flowdir_from_dem_dev(z.data_ptr(), v8.data_ptr(), a.n, a.n, d8.data_ptr())
    del z, v8
else:
    d8 = synth.d8_grid(a.n, a.n, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter(); plan = capi.Plan(dev_ptr=d8.data_ptr(), shape=(a.n, a.n)); t_plan = time.perf_counter() - t0
del d8
plan.set_stream(st.cuda_stream)
import numpy as np, ctypes as C
lo = np.empty(plan.nlevels + 1, np.int32)
capi._check(capi.lib().wfa_plan_export(plan._h, None, None, None, None, None, lo.ctypes.data_as(C.POINTER(C.c_int32)), None))
widths = [int(lo[i + 1] - lo[i]) for i in range(len(lo) - 1)]
print(json.dumps({"n": a.n, "terrain": a.terrain, "widths_at": {str(k): widths[k] for k in (0, 1, 2, 5, 10, 20, 50, 100, 150, 200, 250, 300, 400, 500, 600, 800, 1000, 2000) if k < len(widths)}, "levels": plan.nlevels, "norder": plan.norder, "plan_s": t_plan, "basins": plan.basin_info(), "reaches": plan.reach_info(), "segments": plan.segment_info()}), flush=True)
for ld in [int(x) for x in a.batches.split(",")]:
    A = torch.empty((a.n * a.n, ld), dtype=torch.float64, device=dev)
    for fl in [int(x) for x in a.flags.split(",")]:
        best = 1e30
        for rep in range(a.reps):
            torch.manual_seed(1); A.uniform_(0.0, 100.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.accumulate_dev(A.data_ptr(), ld, ld, fl); e1.record(); e1.synchronize()
            if e0.elapsed_time(e1) < best:
                best, phases = e0.elapsed_time(e1), plan.last_phase_ms()
        if fl == int(a.flags.split(",")[0]):
            ref = int(A.view(torch.int64).sum())        # wrap-around sum of the bit patterns
        else:
            assert int(A.view(torch.int64).sum()) == ref, "sweeps differ"
        units = a.n * a.n * ld
        print(json.dumps({"ld": ld, "flags": fl, "ms": best, "wide_ms": phases[0], "tail_ms": phases[1], "cell_months_per_s": units / best * 1e3,
                          "GBps_algorithmic": units * 16 / best * 1e3 / 1e9}), flush=True)
    del A; torch.cuda.empty_cache()
