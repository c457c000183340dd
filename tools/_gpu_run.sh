#!/bin/bash
export B2_ARGB_FRAMES=64 B2_OPS=predictor_forward
for v in "" p1 p2 p4 p8; do
  echo "== ${v:-base}"
  if [ -n "$v" ]; then export B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so; fi
  python tools/quick_bench.py vp8l 2>&1 | tail -1
done
