import ctypes as C, sys
sys.path.insert(0,'/root/repo')
from gr2msemidistr_b200 import capi
L=capi.lib(); capi.set_device(0)
for mode,name in enumerate(["fma(a,m,const)","dadd(a,m)","dmul(a,m)","fma(a,a,c)"]):
    t=C.c_double()
    rc=L.gr2m_fp64_peak_mix(C.c_int(mode), C.c_int(20000), C.byref(t)); print(name, rc, round(t.value,3), "T thread-instr/s")
print("fma(a,const,const)", capi.fp64_peak_tflops()/2 if hasattr(capi,'fp64_peak_tflops') else None)
print("fma(a,m,c) rrr", capi.fp64_peak_rrr_tflops()/2)
