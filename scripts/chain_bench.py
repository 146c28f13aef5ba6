"""Per-level latency of the upper-phase kernels on synthetic long chains (development aid, GPU only).
64 independent serpentine channels (one basin each), every level exactly one cell wide per basin."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
from gr2msemidistr_b200 import capi
N = -32768
g.build(); capi.set_device(0)
nb, rows, ncol = 64, 16, 513          # per band: 16 rows (8 channel rows), 511-cell runs
band = np.full((rows, ncol), N, np.int16)
for r in range(1, rows - 1, 2):
    if (r // 2) % 2 == 0:
        band[r, 1:ncol - 1] = 1; band[r, ncol - 2] = 7; band[r + 1, ncol - 2] = 7
    else:
        band[r, 1:ncol - 1] = 5; band[r, 1] = 7; band[r + 1, 1] = 7
band[rows - 2:, :] = N
d = np.vstack([band] * nb)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
plan = capi.Plan(d); plan.set_stream(st.cuda_stream)
info = plan.basin_info()
print(json.dumps({"levels": plan.nlevels, "cells": plan.norder, "basins": info}))
for ld in (32, 8):
    A = torch.rand((d.size, ld), dtype=torch.float64, device=dev)
    for fl, name in ((0, "basin+forward"), (16, "basin"), (4, "levelsync")):
        best = 1e9
        for _ in range(3):
            A.uniform_(0, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.accumulate_dev(A.data_ptr(), ld, ld, fl); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"ld": ld, "mode": name, "ms": best, "us_per_level": 1e3 * best / plan.nlevels}))
