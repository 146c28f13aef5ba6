python bench.py --steps 2 --warmup 3 --route-n 0 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('shipped full C4 kernel_ms %.2f' % d['roofline']['kernel_ms'])"
GR2M_CALIB_MAGIC=1 python bench.py --steps 2 --warmup 3 --route-n 0 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('magic=1 full C4 kernel_ms %.2f' % d['roofline']['kernel_ms'])"
GR2M_CALIB_GPS=4 python bench.py --steps 2 --warmup 3 --route-n 0 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gps=4 full C4 kernel_ms %.2f' % d['roofline']['kernel_ms'])"
WHATIFS=7 bash scripts/whatif_full.sh
