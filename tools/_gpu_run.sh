#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "palette" 2>&1 | tail -8
echo "== default (4 stages, 2 CTAs)"; python tools/quick_bench.py palette 2>&1 | tail -4
for v in s3c3 s2c3 s6c2; do echo "== $v"; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py palette 2>&1 | tail -4; done
