timeout 600 python -m pytest tests -m gpu -x -q -k "rescaler or maximum_dimension" 2>&1 | tail -3
timeout 300 python tools/quick_bench.py rescale 2>&1 | tail -3
