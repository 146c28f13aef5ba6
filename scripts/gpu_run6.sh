#!/bin/bash
set -o pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
python -c "
from gr2msemidistr_b200 import capi
capi.set_device(0)
print('fp64 peak (const operands)', capi.fp64_peak_tflops(4096), 'TFLOP/s;  3 register operands', capi.fp64_peak_rrr_tflops(4096), 'TFLOP/s')
" | tee gpurun_out/fp64_peaks.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/route_bench.py --n 20000 --batches 16,32 > gpurun_out/route_20000.log 2>&1; cat gpurun_out/route_20000.log
