#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests/test_gpu_wfa.py tests/test_gpu_golden.py tests/test_gpu_device_api.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python scripts/chain_bench.py 2>&1 | grep "{" | tee gpurun_out/chain.log
timeout 600 python scripts/route_bench.py --n 20000 --batches 32 --flags 0,16 2>&1 | grep "{" | tee gpurun_out/route_20000.log
timeout 600 python scripts/route_bench.py --n 8000 --batches 32,16 --flags 0,16 2>&1 | grep "{" | tee gpurun_out/route_8000.log
