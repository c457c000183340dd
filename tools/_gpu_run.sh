export B2_OPS=predictor_inverse
for f in "" go_webp_b200/lib/variants/*.so; do
echo "== $f"
for nf in 8 64; do B2_LIB=$f B2_ARGB_FRAMES=$nf timeout 120 python tools/quick_bench.py vp8l; done
done
