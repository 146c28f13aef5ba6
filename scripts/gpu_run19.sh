#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --route-n 0 --no-cpu > gpurun_out/bench_nocpu.log 2>&1
tail -1 gpurun_out/bench_nocpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e'])"
