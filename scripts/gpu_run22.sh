#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/route_bench.py --n 8000 --batches 32 --flags 0"
timeout 600 $CMD > gpurun_out/route_8000.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_accum_components" -c 1 -o gpurun_out/route_comp $CMD > gpurun_out/ncu_route.log 2>&1
echo "ncu rc=$?"
