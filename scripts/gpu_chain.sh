#!/bin/bash
mkdir -p gpurun_out
python scripts/chain_bench.py 2>&1 | grep "{" | tee gpurun_out/chain.log
