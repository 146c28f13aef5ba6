python - <<'PY'
import numpy as np
from gr2msemidistr_b200 import capi, synth
capi.set_device(0)
for ncell, nsets in ((64, 8), (128, 8), (256, 8), (1024, 8), (64, 8*148*3)):
    P,E,area,nd=synth.forcing(ncell,480)
    reg=np.zeros(ncell,np.int32)
    par=synth.param_sets(nsets)
    q=synth.qobs_from(np.full(480,100.0))
    with capi.Basin(P,E,area,nd,reg,1) as b:
        ms=[]
        for rep in range(5):
            b.objective_batch(par,q,36,capi.OF_KGE,0,want_crit=False)
            ms.append(b.last_kernel_ms())
    print('ncell',ncell,'nsets',nsets,'kernel ms',[round(m,3) for m in ms])
PY
