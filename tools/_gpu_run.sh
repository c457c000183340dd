timeout 600 python -m pytest tests -m gpu -x -q -k "level_smoothing or maximum_dimension" 2>&1 | tail -2
timeout 300 python tools/quick_bench.py dequant 2>&1 | tail -12
