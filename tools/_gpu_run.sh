timeout 600 python -m pytest tests -m gpu -x -q -k "conversion or convert or sampler or rescaled or output" 2>&1 | tail -2
B2_FRAMES=64 timeout 300 python tools/quick_bench.py convert 2>&1 | grep -i "yuv"
