#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "palette or config0 or config3" 2>&1 | tail -3
python tools/quick_bench.py palette 2>&1 | tail -4
B2_ARGB_FRAMES=64 python tools/quick_bench.py palette 2>&1 | tail -4
