#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_filters or maximum_dimension_next" 2>&1 | tail -2
python tools/quick_bench.py filters 2>&1 | grep -i "unfilter horizontal"
for v in a6 a8 a8c6 a3; do echo $v; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py filters 2>&1 | grep -i "unfilter horizontal"; done
