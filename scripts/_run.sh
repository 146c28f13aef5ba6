python -m pytest tests/test_gpu_d8prep.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -25 | cut -c1-250
