#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "alpha_filters or fullsize" 2>&1 | tail -3
echo "== default 16 warps"; python tools/quick_bench.py filters 2>&1 | grep gradient
for v in g12 g24 g32; do echo "== $v"; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py filters 2>&1 | grep "unfilter gradient"; done
