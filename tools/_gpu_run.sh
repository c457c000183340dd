#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "color_index or vp8l or palette or config2 or maximum" 2>&1 | tail -3 | cut -c1-300
export B2_ARGB_FRAMES=32 B2_OPS=color_index_inverse
for w in 4096 4095; do
  echo "== $w"; B2_ARGB_W=$w python tools/quick_bench.py vp8l 2>&1 | tail -4 | head -1
done
