import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from gr2msemidistr_b200 import capi, synth
capi.set_device(0)
P, E, area, nd = synth.forcing(100000, 480)
region = np.zeros(100000, np.int32)
par = synth.param_sets(10000)
qobs = synth.qobs_from(np.full(480, 40.0))
Pp, Ep = torch.from_numpy(P).pin_memory(), torch.from_numpy(E).pin_memory()
for nsets in (10000, 10000, 10000):
    t0 = time.perf_counter(); b = capi.Basin(Pp.numpy(), Ep.numpy(), area, nd, region, 1); t1 = time.perf_counter()
    r = b.objective_batch(par[:nsets], qobs, 36, capi.OF_KGE, 0, want_crit=False); t2 = time.perf_counter()
    torch.cuda.synchronize(); t2b = time.perf_counter()
    b.close(); t3 = time.perf_counter()
    print(nsets, "create %.4f obj %.4f sync %.4f destroy %.4f" % (t1 - t0, t2 - t1, t2b - t2, t3 - t2b), flush=True)
