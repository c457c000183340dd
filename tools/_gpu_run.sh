#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_filters or maximum_dimension_next or rescale" 2>&1 | tail -3
python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"
