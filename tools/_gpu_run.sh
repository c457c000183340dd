#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_multipl or sampler or convert" 2>&1 | tail -3 | cut -c1-200
python tools/quick_bench.py convert 2>&1 | grep -i "YUV\|point"
