#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "predictor or vp8l or chain or config2 or resident" 2>&1 | tail -4
export B2_OPS=predictor_inverse
for n in 1 64 128; do B2_ARGB_FRAMES=$n python tools/quick_bench.py vp8l 2>&1 | tail -1; done
echo u8; for n in 1 64; do B2_LIB=go_webp_b200/lib/variants/libb2webp_u8.so B2_ARGB_FRAMES=$n python tools/quick_bench.py vp8l 2>&1 | tail -1; done
