#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_filters or maximum_dimension_next" 2>&1 | tail -3
python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"
for v in w12 w8; do echo $v; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"; done
B2_FRAMES=296 python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"
