timeout 600 python -m pytest tests -m gpu -x -q -k "import or mixed or config1 or maximum" 2>&1 | tail -2
B2_MIXED_BREAKDOWN=1 timeout 600 python tools/mixed_bench.py 2>&1 | tail -4 | cut -c1-330
B2_FRAMES=256 python tools/quick_bench.py import | head -2
