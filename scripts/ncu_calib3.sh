# ncu --set full capture of the recursion kernel on the 1.2e10-unit slice (run only after the plain command exited 0)
SMALL="python bench.py --steps 2 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
OUT=${OUT:-calib4}
$SMALL > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib -s 3 -c 1 -o gpurun_out/$OUT -f $SMALL > gpurun_out/ncu_c3.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_c3.log | cut -c1-200
