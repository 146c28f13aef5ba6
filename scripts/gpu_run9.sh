#!/bin/bash
set -o pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
SMALL="python bench.py --steps 3 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
timeout 600 $SMALL > gpurun_out/bench_small.log 2>&1
tail -1 gpurun_out/bench_small.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['clocks'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib -s 3 -c 1 -o gpurun_out/calib_r1d $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
