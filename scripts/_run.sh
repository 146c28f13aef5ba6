python - <<'PY'
import time, numpy as np
from gr2msemidistr_b200 import capi, synth
capi.set_device(0)
P,E,area,nd=synth.forcing(100000,480)
reg=np.zeros(100000,np.int32)
par=np.array(synth.SET0)
for rep in range(4):
    t0=time.perf_counter()
    with capi.Basin(P,E,area,nd,reg,1) as b:
        t1=time.perf_counter()
        out=b.run(par)
        t2=time.perf_counter()
    print('create %.3f s  run %.3f s  run e2e %.3g cell-months/s'%(t1-t0,t2-t1,4.8e7/(t2-t1)))
PY
