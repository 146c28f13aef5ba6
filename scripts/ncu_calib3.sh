SMALL="python bench.py --steps 2 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
export GR2M_CALIB3_CTAS=3
$SMALL > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib3 -s 3 -c 1 -o gpurun_out/calib3_v2 -f $SMALL > gpurun_out/ncu_c3.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_c3.log
