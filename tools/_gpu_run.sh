#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "predictor or vp8l or chain or config2" 2>&1 | tail -3
for r in 1 2; do
echo "== fast"; B2_ARGB_FRAMES=64 B2_OPS=predictor_inverse python tools/quick_bench.py vp8l 2>&1 | tail -1
echo "== fast pub8"; B2_LIB=go_webp_b200/lib/variants/libb2webp_pub8.so B2_ARGB_FRAMES=64 B2_OPS=predictor_inverse python tools/quick_bench.py vp8l 2>&1 | tail -1
done
