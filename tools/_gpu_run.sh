timeout 600 python -m pytest tests -m gpu -x -q -k "vp8l" 2>&1 | tail -3
