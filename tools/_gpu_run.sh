#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "predictor or vp8l or dither or chain or staged" 2>&1 | tail -15
python tools/quick_bench.py vp8l 2>&1 | grep -i pred
