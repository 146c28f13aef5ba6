#!/bin/bash
set -o pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e > gpurun_out/bench_small.log 2>&1
tail -1 gpurun_out/bench_small.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
