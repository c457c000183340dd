#!/bin/bash
python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"
for v in w12 w14 s500 s1000; do echo $v; B2_LIB=go_webp_b200/lib/variants/libb2webp_$v.so python tools/quick_bench.py filters 2>&1 | grep -i "unfilter gradient"; done
