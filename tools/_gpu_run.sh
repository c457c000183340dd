#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "picture_tools or cleanup or blend" 2>&1 | tail -3
B2_FRAMES=128 python tools/quick_bench.py tools 2>&1 | tail -8
