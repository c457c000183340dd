timeout 300 python -m pytest tests -m gpu -x -q -k "alpha_filters or maximum_dimension" 2>&1 | tail -2
timeout 300 python tools/quick_bench.py filters 2>&1 | grep vertical
