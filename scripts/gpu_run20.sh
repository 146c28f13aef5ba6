#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
SMALL="python bench.py --steps 2 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
timeout 600 $SMALL > gpurun_out/bench_small.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib -s 3 -c 1 -o gpurun_out/calib_r1e $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
