#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python examples/bundled_example.py > gpurun_out/example.log 2>&1; echo "example rc=$?"; tail -5 gpurun_out/example.log | cut -c1-300
