timeout 600 python -m pytest tests -m gpu -x -q -k "alpha_filters" 2>&1 | tail -3
timeout 600 python tools/quick_bench.py filters 2>&1 | tail -8
