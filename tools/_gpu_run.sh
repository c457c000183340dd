#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "predictor or vp8l_golden or config2 or maximum_dimension" 2>&1 | tail -5 | cut -c1-300
export B2_ARGB_FRAMES=32 B2_OPS=predictor_forward
for w in 4096 4095 4094; do
  B2_ARGB_W=$w python tools/quick_bench.py vp8l 2>&1 | tail -1
done
