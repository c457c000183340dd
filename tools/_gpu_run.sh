#!/bin/bash
for v in r8c4 r2c8 r4c6 r6c5; do echo $v; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py filters 2>&1 | grep -i "unfilter horizontal"; done
