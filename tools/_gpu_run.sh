export B2_OPS=predictor_forward B2_ARGB_FRAMES=64
for f in "" go_webp_b200/lib/variants/*.so; do echo "== $f"; B2_LIB=$f timeout 120 python tools/quick_bench.py vp8l | tail -1; done
