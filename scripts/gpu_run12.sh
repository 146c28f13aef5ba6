#!/bin/bash
# 2-GPU scaling check: torchrun launch exactly as the driver does it
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"
tail -1 gpurun_out/bench_n2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref rc=$?"
tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
