#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_filters or filter" 2>&1 | tail -3 | cut -c1-200
python tools/quick_bench.py filters 2>&1 | grep -i "filter  "
