#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1200 python -m pytest tests/test_gpu_wfa.py tests/test_gpu_golden.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for w in 4096 16384 65536; do WFA_COMP_WIDTH=$w timeout 600 python scripts/route_bench.py --n 20000 --batches 32 --flags 0 2>&1 | grep "{" ; done | tee gpurun_out/route_20000.log
