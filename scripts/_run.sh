python -m pytest tests/test_r_shim.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -8 | cut -c1-250
