#!/bin/bash
mkdir -p gpurun_out
python scripts/chain_bench.py > gpurun_out/chain_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_accum_components_fw -c 1 -o gpurun_out/chain_comp python scripts/chain_bench.py > gpurun_out/ncu_chain.log 2>&1
echo "rc=$?"
