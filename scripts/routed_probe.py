"""Routed calibration probe: bench.py's routed_calibration figure alone (grid side N, default 6000), also for ncu launch lists:
    python scripts/routed_probe.py 6000 && ncu --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/routed_launches.csv python scripts/routed_probe.py 6000"""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gr2msemidistr_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
nsets = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
capi.set_device(0)
torch.cuda.set_device(0)
print(json.dumps(bench.bench_routed_calibration(types.SimpleNamespace(months=480), capi, torch, n=n, nsets=nsets)))
