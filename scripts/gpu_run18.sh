#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default rc=$?"
tail -1 gpurun_out/bench_default.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"
FULL="python bench.py --steps 1 --warmup 3 --route-n 0 --no-cpu --no-e2e"
timeout 600 $FULL > gpurun_out/bench_full_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib2 -s 3 -c 1 -o gpurun_out/calib_full_c4 $FULL > gpurun_out/ncu_full_c4.log 2>&1
echo "ncu rc=$?"
