#!/bin/bash
# Recursion-kernel experiment: parity tests, then the 1.2e10-unit slice with each kernel variant.
#   gpurun --timeout 1200 -- 'bash scripts/calib_variants.sh'
# variants: "generic" or "<CTAs per SM>:<cells per thread>:<rounding form: 0 FRND+F2I, 1 magic add, 2 fixed point, 3 fixed point + clamp-free pairs>" of the table-mode kernel
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_gr2m.py tests/test_gpu_golden.py tests/test_reference_pin.py -m gpu -q -x > gpurun_out/pytest_calib.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_calib.log
fi
SMALL="python bench.py --steps 3 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
for v in ${VARIANTS:-generic 3:2:0 3:2:2 3:2:3}; do
  if [ $v = generic ]; then export GR2M_CALIB_GENERIC=1; else unset GR2M_CALIB_GENERIC; IFS=: read c n m <<< "$v"; export GR2M_CALIB_CTAS=$c GR2M_CALIB_CPT=$n GR2M_CALIB_MAGIC=$m; fi
  $SMALL > gpurun_out/slice_$v.log 2>&1; rc=$?
  tail -1 gpurun_out/slice_$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant $v rc=$rc  value %.4g  kernel_ms %.2f  checksum %s' % (d['value'], d['roofline']['kernel_ms'], d.get('objective_checksum')))"
done
