#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests/test_gpu_gr2m.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default rc=$?"
tail -1 gpurun_out/bench_default.log | cut -c1-200
CMD="python bench.py --steps 2 --warmup 1 --route-n 0 --no-cpu --no-e2e --sets 2500 --cells 20000"
timeout 600 $CMD > gpurun_out/bench_hl.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_headline.csv $CMD > gpurun_out/ncu_ll.log 2>&1
echo "ncu rc=$?"
