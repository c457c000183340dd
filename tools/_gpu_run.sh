timeout 600 python -m pytest tests -m gpu -x -q -k "picture_tools or maximum_dimension" 2>&1 | tail -2
B2_FRAMES=128 timeout 300 python tools/quick_bench.py tools 2>&1 | grep -i "cleanup"
