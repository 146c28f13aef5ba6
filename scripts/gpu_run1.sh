#!/bin/bash
# first GPU pass: parity tests, small + full bench, ncu launch list of the small bench
set -o pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
SMALL="python bench.py --steps 2 --warmup 1 --sets 512 --cells 20000 --route-n 3000 --no-cpu"
timeout 600 $SMALL > gpurun_out/bench_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_small.csv $SMALL > gpurun_out/ncu_small.log 2>&1
echo "small rc=$?"
tail -2 gpurun_out/bench_small.log
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "full rc=$?"
tail -2 gpurun_out/bench_full.log
