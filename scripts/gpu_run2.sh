#!/bin/bash
set -o pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python scripts/route_bench.py --n 8000 > gpurun_out/route_8000.log 2>&1; cat gpurun_out/route_8000.log
timeout 600 python scripts/route_bench.py --n 20000 --batches 16,32 > gpurun_out/route_20000.log 2>&1; cat gpurun_out/route_20000.log
SMALL="python bench.py --steps 2 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
timeout 600 $SMALL > gpurun_out/bench_small.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib -s 3 -c 1 -o gpurun_out/calib_r1 $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/bench_small.log | cut -c1-600
