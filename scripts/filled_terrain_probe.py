"""How long does the device pit fill + D8 take on the survey-spec fractal terrain, and what does its level structure look like?"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
from gr2msemidistr_b200 import capi, synth
g.build(); capi.set_device(0)
dev = torch.device("cuda", 0)
relief = float(os.environ.get("RELIEF", "1.0"))
for n in [int(x) for x in sys.argv[1:]] or [2000, 4000]:
    z, valid = synth.fractal_dem(n, n, device=dev, relief=relief)
    zh, vh = z.cpu().numpy(), valid.cpu().numpy()
    del z, valid
    t0 = time.perf_counter()
    fd, filled = capi.flowdir_from_dem(zh, vh, want_filled=True)
    t1 = time.perf_counter()
    lake = float(np.mean((filled > zh + 1e-9)[vh]))
    with capi.Plan(fd) as plan:
        info = {"n": n, "relief": relief, "flowdir_s": t1 - t0, "lake_share": lake, "levels": plan.nlevels, "norder": plan.norder,
                "launches": capi.kernel_launches(), "basins": plan.basin_info(), "reaches": plan.reach_info()}
    print(json.dumps(info), flush=True)
