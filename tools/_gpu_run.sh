#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "rescale" 2>&1 | tail -3 | cut -c1-200
python tools/quick_bench.py rescale 2>&1 | tail -4
