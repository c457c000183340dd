timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B2_FRAMES=64 timeout 300 python tools/quick_bench.py convert 2>&1 | tail -13
