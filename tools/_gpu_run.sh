#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "rescale" 2>&1 | tail -3
python tools/quick_bench.py rescale 2>&1 | tail -3
B2_RESCALE_NOFAST=1 python tools/quick_bench.py rescale 2>&1 | tail -3
