#!/bin/bash
# Recursion-kernel experiment: parity tests, then the 1.2e10-unit slice with each kernel variant.
#   gpurun --timeout 1200 -- 'bash scripts/calib_variants.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gr2m.py tests/test_gpu_golden.py tests/test_reference_pin.py -m gpu -q -x > gpurun_out/pytest_calib.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_calib.log
SMALL="python bench.py --steps 3 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
for v in generic 2 3 4; do
  if [ $v = generic ]; then export GR2M_CALIB_GENERIC=1; else unset GR2M_CALIB_GENERIC; export GR2M_CALIB3_CTAS=$v; fi
  $SMALL > gpurun_out/slice_$v.log 2>&1; echo "variant $v rc=$?"
  tail -1 gpurun_out/slice_$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  value %.4g  kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))"
done
