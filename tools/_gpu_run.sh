#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "rescale" 2>&1 | tail -3
python tools/quick_bench.py rescale 2>&1 | tail -3
for v in rsf8 rsf4; do echo $v; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py rescale 2>&1 | tail -3; done
