python -m pytest tests/test_gpu_wfa.py tests/test_gpu_router.py tests/test_gpu_guard.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-250
python scripts/route_bench.py --n 8000 --batches 32 --flags 0,256,128 2>&1 | tail -3 | cut -c1-200
