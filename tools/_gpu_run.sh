#!/bin/bash
for v in f8c3 f8c4 f6c4 f4c6 f2c8; do echo $v; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py filters 2>&1 | grep -i "alpha filter"; done
