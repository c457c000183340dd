timeout 300 python -m pytest tests -m gpu -x -q -k "alpha_filters or maximum_dimension or golden" 2>&1 | tail -3
timeout 120 python tools/quick_bench.py filters 2>&1 | grep gradient
