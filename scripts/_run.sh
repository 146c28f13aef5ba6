echo "--- guard on"; GR2M_GUARD=1 python scripts/_dbg.py 2>&1 | tail -5
for i in 1 2 3; do python -m pytest tests/test_gpu_guard.py -m gpu -q -x 2>&1 | tail -2; done
