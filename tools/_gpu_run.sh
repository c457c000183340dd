export B2_OPS=predictor_inverse B2_ARGB_FRAMES=64
for f in "" go_webp_b200/lib/variants/*.so; do echo "== $f"; B2_LIB=$f timeout 120 python tools/quick_bench.py vp8l | tail -1; done
B2_LIB=go_webp_b200/lib/variants/libb2webp_c8p8.so timeout 300 python -m pytest tests -m gpu -x -q -k "predict" 2>&1 | tail -2
