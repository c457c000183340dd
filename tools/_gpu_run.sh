timeout 300 python -m pytest tests -m gpu -x -q -k "predict or config2" 2>&1 | tail -3
export B2_OPS=predictor_inverse
B2_ARGB_FRAMES=1 B2_ARGB_H=32 timeout 120 python tools/quick_bench.py vp8l
B2_ARGB_FRAMES=1 timeout 120 python tools/quick_bench.py vp8l
for nf in 8 64; do B2_ARGB_FRAMES=$nf timeout 120 python tools/quick_bench.py vp8l; done
