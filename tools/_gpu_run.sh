#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "smoothing or maximum_dimension_next or rescaled" 2>&1 | tail -3
python tools/quick_bench.py dequant 2>&1 | tail -12
