#!/bin/bash
# Everything the round-end driver runs, in one gpurun call:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh'
set -o pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-160
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?"
# ncu recipes used for profiles/ (one ncu command per call, only after the plain run above exited 0):
#   SMALL="python bench.py --steps 2 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
#   $SMALL && ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib -s 3 -c 1 -o gpurun_out/calib $SMALL
#   python scripts/ncu_summary.py gpurun_out/calib.ncu-rep 1.2e10 > profiles/<name>.txt
