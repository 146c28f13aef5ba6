python -m pytest tests/test_gpu_routed_objective.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -12 | cut -c1-250
