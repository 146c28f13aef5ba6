#!/bin/bash
# Timing sensitivities of the table-mode recursion kernel: rebuild with one instruction group dropped
# (GR2M_WHATIF=k, results are wrong on purpose) and time the 1.2e10-unit slice.  See gr2m_calib3.cuh.
mkdir -p gpurun_out /tmp/wi
SMALL="python bench.py --steps 3 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
C=gr2msemidistr_b200/csrc
for k in ${WHATIFS:-0 1 5 7}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DGR2M_WHATIF=$k -o /tmp/wi/lib$k.so $C/gr2m.cu $C/wfa.cu $C/d8prep.cu $C/router.cu 2>/dev/null || { echo "build $k failed"; continue; }
  GR2M_B200_LIB=/tmp/wi/lib$k.so $SMALL > gpurun_out/whatif_$k.log 2>&1
  tail -1 gpurun_out/whatif_$k.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('whatif $k  kernel_ms %.2f' % d['roofline']['kernel_ms'])"
done
