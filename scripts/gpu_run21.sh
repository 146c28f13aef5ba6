#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
CMD="python scripts/route_bench.py --n 8000 --batches 32 --flags 0"
timeout 600 $CMD > gpurun_out/route_8000.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_accum_level|k_accum_components" -c 4 -o gpurun_out/route_r1 $CMD > gpurun_out/ncu_route.log 2>&1
echo "ncu rc=$?"; grep "{" gpurun_out/route_8000.log
