#!/bin/bash
# A/B of a -DGR2M_WHATIF=k build against the shipped one on the FULL C4 launch
mkdir -p gpurun_out /tmp/wi
C=gr2msemidistr_b200/csrc
for k in ${WHATIFS:-0 8}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DGR2M_WHATIF=$k -o /tmp/wi/lib$k.so $C/gr2m.cu $C/wfa.cu $C/d8prep.cu $C/router.cu 2>/dev/null || { echo "build $k failed"; continue; }
  GR2M_B200_LIB=/tmp/wi/lib$k.so python bench.py --steps 2 --warmup 3 --route-n 0 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('whatif $k full C4 kernel_ms %.2f checksum %s' % (d['roofline']['kernel_ms'], d['objective_checksum']))"
done
