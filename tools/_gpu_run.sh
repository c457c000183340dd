timeout 600 python -m pytest tests -m gpu -x -q -k "rescaler" 2>&1 | tail -12
timeout 300 python tools/quick_bench.py rescale 2>&1 | tail -4
