#!/bin/bash
# Time the 1.2e10-unit slice of the recursion kernel with each experiment library in build/exp (see exp_build.sh):
#   gpurun --timeout 900 -- 'bash scripts/exp_run.sh base nst3 ...'   (FULL=name also runs the full C4 launch with it)
mkdir -p gpurun_out
SMALL="python bench.py --steps 3 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
for n in "$@"; do
  GR2M_B200_LIB=$PWD/build/exp/lib$n.so $SMALL > gpurun_out/exp_$n.log 2>&1; rc=$?
  tail -1 gpurun_out/exp_$n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('exp $n rc=$rc  value %.4g  kernel_ms %.2f  checksum %s' % (d['value'], d['roofline']['kernel_ms'], d.get('objective_checksum')))"
done
if [ -n "$FULL" ]; then
  GR2M_B200_LIB=$PWD/build/exp/lib$FULL.so python bench.py --steps 3 --warmup 3 --route-n 0 --no-cpu --no-e2e > gpurun_out/exp_full_$FULL.log 2>&1; rc=$?
  tail -1 gpurun_out/exp_full_$FULL.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('full $FULL rc=$rc  value %.4g  kernel_ms %.2f  frac %.4f checksum %s' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('objective_checksum')))"
fi
