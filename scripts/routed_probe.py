"""Routed calibration probe: the bench's routed_calibration figure alone (grid side N, default 6000), for ncu launch lists:
    python scripts/routed_probe.py 6000 && ncu --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/routed_launches.csv python scripts/routed_probe.py 6000"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gr2msemidistr_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
nsets = int(sys.argv[2]) if len(sys.argv) > 2 else 256
capi.set_device(0)
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
d8 = synth.d8_grid(n, n, device=dev)
plan = capi.Plan(dev_ptr=d8.data_ptr(), shape=(n, n))
plan.set_stream(torch.cuda.current_stream().cuda_stream)
nb = max(1, min(100, n // 40))
src, off, cells = synth.block_polygons(n, n, nb, nb, margin=max(1, n // nb // 4))
router = plan.router(src, off, cells)
print(json.dumps(bench.bench_routed_calibration(None, capi, router, src.shape[0], 480, nsets)))
router.close()
plan.close()
