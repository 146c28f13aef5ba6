python -m pytest tests/test_gpu_gr2m.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"
tail -1 gpurun_out/bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=2 value %.4g e2e %.4g routing %.4g filled %.4g launches %s checksum %s'%(d['value'],d['e2e']['value'],d['routing']['value'],d['routing_filled_terrain']['value'],d['gpu_launches'],d['objective_checksum']))"
