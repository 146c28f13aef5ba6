#!/bin/bash
# compute-sanitizer, one tool per call: bash scripts/gpu_sanitize.sh memcheck|racecheck
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 compute-sanitizer --tool ${1:-memcheck} --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$1.log 2>&1
echo "sanitizer $1 rc=$?"; tail -6 gpurun_out/sanitizer_$1.log
