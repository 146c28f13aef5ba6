#!/bin/bash
# Experiment builds of the library (run here, before gpurun): scripts/exp_build.sh name "-DFLAG ..." [name flags ...]
# -> build/exp/lib<name>.so (build/ is git-ignored but travels to the GPU box); used through GR2M_B200_LIB.
C=gr2msemidistr_b200/csrc
mkdir -p build/exp
[ -f build/exp/rest.o ] || true
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $f -o build/exp/lib$n.so $C/gr2m.cu $C/wfa.cu $C/d8prep.cu $C/router.cu 2>/dev/null &
done
wait
ls -la build/exp
