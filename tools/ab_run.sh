python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/stage_bench.py --iters 5 2>&1 | tail -1
