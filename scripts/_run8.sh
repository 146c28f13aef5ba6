python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --route-n 0 --no-cpu > gpurun_out/bench_n8b.log 2>&1; echo "n8 rc=$?"
tail -1 gpurun_out/bench_n8b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=8 value %.4g ms %.2f e2e %.4g checksum %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['objective_checksum']))"
