#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for w in 1024 4096 32768 131072; do
  echo "COMP_WIDTH=$w"; WFA_COMP_WIDTH=$w timeout 600 python scripts/route_bench.py --n 20000 --batches 32 --flags 0 2>&1 | grep -v Traceback | grep "{"
done | tee gpurun_out/route_compwidth.log
