#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_filters or maximum_dimension_next" 2>&1 | tail -3
python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"
B2_FRAMES=16 python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"
B2_FRAMES=296 python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"
