for t in tilted filled; do python scripts/route_bench.py --n 20000 --batches 32 --flags 0 --terrain $t 2>&1 | tail -1 | cut -c1-200; done
