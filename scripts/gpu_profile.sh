#!/bin/bash
# Round-2 evidence run: default bench, then (only after each plain command exited 0) the ncu passes profiles/ quotes.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh'
mkdir -p gpurun_out
echo "(default bench: see gpu_check.sh)"
SMALL="python bench.py --steps 2 --warmup 3 --sets 1250 --cells 20000 --route-n 0 --no-cpu --no-e2e"
$SMALL > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gr2m_calib -s 3 -c 1 -o gpurun_out/r2_calib4 -f $SMALL > gpurun_out/ncu_c4.log 2>&1; echo "ncu slice rc=$?"
FULL="python bench.py --steps 1 --warmup 3 --route-n 0 --no-cpu --no-e2e"
$FULL > gpurun_out/ncu_plain_full.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_gr2m_calib -s 3 -c 1 --csv --log-file gpurun_out/r2_calib4_full_traffic.csv $FULL > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
HEAD="python bench.py --steps 2 --warmup 3 --no-cpu"
$HEAD > gpurun_out/ncu_plain_head.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_headline.csv $HEAD > gpurun_out/ncu_head.log 2>&1; echo "ncu launches rc=$?"
