timeout 600 python -m pytest tests -m gpu -x -q -k "rescale" 2>&1 | tail -2
B2_FRAMES=64 timeout 300 python tools/quick_bench.py rescale 2>&1 | tail -3
